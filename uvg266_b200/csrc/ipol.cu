// ipol.cu - motion-compensation sample interpolation (next row 8(f)-1) and the fused fractional
// motion-search scoring kernel.
//
// Restates generic/ipol-generic.c:134-210 (uvg_sample_quarterpel_luma(_hi): 8 taps, 1/16 phases) and
// :681-759 (uvg_sample_octpel_chroma(_hi): 4 taps, 1/32 phases): horizontal pass into int16
// (>> bitdepth-8), vertical pass (>> 6), then the uni-prediction rounding to a pixel
// ((v + offset) >> (14 - bitdepth), clipped) or, for the "_hi" variants, the 14-bit intermediate.
// Out-of-frame samples are the replicated border, which is what uvg_get_extended_block builds for
// the encoder (ipol-generic.c:761-811); here it is a clamp of the read coordinates.
//
// Kernels: one CTA per block. The (w + taps - 1) x (h + taps - 1) window is staged once in shared
// memory, the horizontal pass writes int16 rows next to it, the vertical pass reads those.
// The fused kernel (uvg_b200_frac_search_dev) stages a window one sample larger on every side, and
// for each of the k fractional candidates around the block's motion vector interpolates into shared
// memory and scores SATD against the original with the packed Hadamard of hadamard.cuh - the
// interpolated blocks never reach HBM (search_inter.c:1085-1250 interpolates into 4+4 block buffers
// per step and scores them with satd_any_size_quad).
#include "common.cuh"
#include <mutex>
#include "hadamard.cuh"

namespace uvgb {

// VVC interpolation filters (filter.c:62-116). Luma: phases 0..8 of 16, phases 9..15 mirror 7..1.
// Chroma: 32 phases; the same table as the intra cubic filter.
__constant__ signed char c_luma_half[9][8] = {
  { 0, 0, 0, 64, 0, 0, 0, 0 },     { 0, 1, -3, 63, 4, -2, 1, 0 },    { -1, 2, -5, 62, 8, -3, 1, 0 },
  { -1, 3, -8, 60, 13, -4, 1, 0 }, { -1, 4, -10, 58, 17, -5, 1, 0 }, { -1, 4, -11, 52, 26, -8, 3, -1 },
  { -1, 3, -9, 47, 31, -10, 4, -1 }, { -1, 4, -11, 45, 34, -10, 4, -1 }, { -1, 4, -11, 40, 40, -11, 4, -1 } };
__constant__ signed char c_chroma[32][4] = {
  { 0, 64, 0, 0 }, { -1, 63, 2, 0 }, { -2, 62, 4, 0 }, { -2, 60, 7, -1 }, { -2, 58, 10, -2 }, { -3, 57, 12, -2 },
  { -4, 56, 14, -2 }, { -4, 55, 15, -2 }, { -4, 54, 16, -2 }, { -5, 53, 18, -2 }, { -6, 52, 20, -2 },
  { -6, 49, 24, -3 }, { -6, 46, 28, -4 }, { -5, 44, 29, -4 }, { -4, 42, 30, -4 }, { -4, 39, 33, -4 },
  { -4, 36, 36, -4 }, { -4, 33, 39, -4 }, { -4, 30, 42, -4 }, { -4, 29, 44, -5 }, { -4, 28, 46, -6 },
  { -3, 24, 49, -6 }, { -2, 20, 52, -6 }, { -2, 18, 53, -5 }, { -2, 16, 54, -4 }, { -2, 15, 55, -4 },
  { -2, 14, 56, -4 }, { -2, 12, 57, -3 }, { -2, 10, 58, -2 }, { -1, 7, 60, -2 }, { 0, 4, 62, -2 }, { 0, 2, 63, -1 } };

template <int TAPS>
__device__ __forceinline__ void load_filter(int phase, int (&f)[TAPS])
{
  if (TAPS == 8) {
#pragma unroll
    for (int i = 0; i < TAPS; ++i) f[i] = phase <= 8 ? c_luma_half[phase][i] : c_luma_half[16 - phase][7 - i];
  } else {
#pragma unroll
    for (int i = 0; i < TAPS; ++i) f[i] = c_chroma[phase][i];
  }
}

struct IpolArgs {
  const void *plane; int stride, plane_w, plane_h;      // reference picture plane (pixels)
  int w, h;                                              // block
  const int32_t *bx, *by, *mvx, *mvy;                    // per block: origin in the plane, MV in filter units
  void *dst;                                             // dense [n][h*w] pixels (or int16 when hi)
  int hi, shift1, wp_shift, pixel_max;
};

// ---- plain interpolation: one CTA per block -----------------------------------------------------
template <typename Pix, int TAPS>
__global__ void __launch_bounds__(128)
ipol_kernel(IpolArgs a, size_t n_blocks)
{
  constexpr int OFF = TAPS == 8 ? 3 : 1, PAD = TAPS - 1, FRAC = TAPS == 8 ? 4 : 5, MASK = (1 << FRAC) - 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int w = a.w, h = a.h, ww = w + PAD, wh = h + PAD;
  Pix *win = reinterpret_cast<Pix *>(smem_raw);                                  // [wh][ww]
  int16_t *hor = reinterpret_cast<int16_t *>(smem_raw + (((size_t)ww * wh * sizeof(Pix) + 15) & ~(size_t)15));   // [wh][w]
  const Pix *plane = static_cast<const Pix *>(a.plane);
  for (size_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
    const int mvx = a.mvx[b], mvy = a.mvy[b];
    const int x0 = a.bx[b] + (mvx >> FRAC) - OFF, y0 = a.by[b] + (mvy >> FRAC) - OFF;
    int fh[TAPS], fv[TAPS];
    load_filter<TAPS>(mvx & MASK, fh); load_filter<TAPS>(mvy & MASK, fv);
    __syncthreads();
    for (int i = threadIdx.x; i < ww * wh; i += blockDim.x) {
      const int r = i / ww, c = i - r * ww;
      const int yy = min(max(y0 + r, 0), a.plane_h - 1), xx = min(max(x0 + c, 0), a.plane_w - 1);
      win[i] = plane[(size_t)yy * a.stride + xx];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < wh * w; i += blockDim.x) {
      const int r = i / w, c = i - r * w;
      int s = 0;
#pragma unroll
      for (int t = 0; t < TAPS; ++t) s += fh[t] * (int)win[r * ww + c + t];
      hor[i] = (int16_t)(s >> a.shift1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
      const int r = i / w, c = i - r * w;
      int s = 0;
#pragma unroll
      for (int t = 0; t < TAPS; ++t) s += fv[t] * (int)hor[(r + t) * w + c];
      s >>= 6;
      if (a.hi) static_cast<int16_t *>(a.dst)[b * (size_t)(w * h) + i] = (int16_t)s;
      else static_cast<Pix *>(a.dst)[b * (size_t)(w * h) + i] =
          (Pix)min(max((s + (1 << (a.wp_shift - 1))) >> a.wp_shift, 0), a.pixel_max);
    }
  }
}

// ---- fused fractional search: interpolate k candidates around the MV and score SATD, 8-bit luma ----
// Per block: (1) the window covering every candidate (integer offsets -1..+1 around the MV, filter
// reach 3 / 4) is staged once; (2) the horizontal pass runs once per DISTINCT horizontal position of
// the candidate set (5 for the 17-point half/quarter-pel pattern, like filter_hpel/qpel_blocks do),
// 4 outputs per step with dp4a on the window bytes; (3) one thread per (candidate, 8x8 sub-block)
// runs the vertical pass for its 64 pixels in registers and feeds the packed Hadamard directly.
struct FracArgs {
  const uint8_t *cur, *ref; int stride, plane_w, plane_h;
  int h;                                                 // multiple of 8, <= 64
  const int32_t *bx, *by, *mvx, *mvy;                    // block origin, base MV in 1/16 samples
  int k, nx;
  signed char gdx[8];                                    // distinct horizontal offsets (1/16 samples)
  signed char cg[64], cdy[64];                           // per candidate: index into gdx, vertical offset
  uint32_t *out;                                         // [n][k] satd_any_size(orig, interpolated)
};

__device__ __forceinline__ int dp4a_us8(uint32_t a, int b, int c)
{
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

template <int W, int G>      // W: block width; G: blocks per CTA
__global__ void __launch_bounds__(128)
frac_search_kernel(FracArgs a, size_t n_blocks)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int h = a.h, wh = h + 10;
  constexpr int P = (W + 10 + 3 + 3) & ~3;               // window row pitch (words aligned, 3 spare bytes for the last word)
  const int win_bytes = (wh * P + 15) & ~15, hor_elems = a.nx * wh * W;
  const int slot_bytes = win_bytes + W * h + ((hor_elems * 2 + 15) & ~15);
  const int nsx = W / 8, nsub = nsx * (h / 8);
  for (size_t b0 = (size_t)blockIdx.x * G; b0 < n_blocks; b0 += (size_t)gridDim.x * G) {
    __syncthreads();
    // (1) windows and originals of the CTA's G blocks
    for (int i = threadIdx.x; i < G * wh * (W + 10); i += blockDim.x) {
      const int g = i / (wh * (W + 10)), rem = i - g * (wh * (W + 10)), r = rem / (W + 10), c = rem - r * (W + 10);
      const size_t b = min(b0 + g, n_blocks - 1);
      const int x0 = a.bx[b] + (a.mvx[b] >> 4) - 4, y0 = a.by[b] + (a.mvy[b] >> 4) - 4;
      const int yy = min(max(y0 + r, 0), a.plane_h - 1), xx = min(max(x0 + c, 0), a.plane_w - 1);
      smem_raw[g * slot_bytes + r * P + c] = a.ref[(size_t)yy * a.stride + xx];
    }
    for (int i = threadIdx.x; i < G * W * h; i += blockDim.x) {
      const int g = i / (W * h), rem = i - g * (W * h), r = rem / W, c = rem - r * W;
      const size_t b = min(b0 + g, n_blocks - 1);
      smem_raw[g * slot_bytes + win_bytes + rem] =
          a.cur[(size_t)min(a.by[b] + r, a.plane_h - 1) * a.stride + min(a.bx[b] + c, a.plane_w - 1)];
    }
    __syncthreads();
    // (2) horizontal pass per distinct horizontal position: item = (block, group, row, 4 columns)
    for (int i = threadIdx.x; i < G * a.nx * wh * (W / 4); i += blockDim.x) {
      const int g = i / (a.nx * wh * (W / 4)), rem = i - g * (a.nx * wh * (W / 4));
      const int gx = rem / (wh * (W / 4)), rem2 = rem - gx * (wh * (W / 4)), r = rem2 / (W / 4), c0 = 4 * (rem2 - r * (W / 4));
      const size_t b = min(b0 + g, n_blocks - 1);
      const int mvx = a.mvx[b], cx = mvx + a.gdx[gx], relx = (cx >> 4) - (mvx >> 4);
      int f[8];
      load_filter<8>(cx & 15, f);
      const int flo = (f[0] & 255) | ((f[1] & 255) << 8) | ((f[2] & 255) << 16) | ((f[3] & 255) << 24);
      const int fhi = (f[4] & 255) | ((f[5] & 255) << 8) | ((f[6] & 255) << 16) | ((f[7] & 255) << 24);
      const unsigned char *row = smem_raw + g * slot_bytes + r * P;
      const int j0 = 1 + relx + c0, al = j0 & ~3, sh = (j0 & 3) * 8;
      const uint32_t *wp = reinterpret_cast<const uint32_t *>(row + al);
      const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
      const uint32_t A0 = __funnelshift_r(w0, w1, sh), A1 = __funnelshift_r(w1, w2, sh), A2 = __funnelshift_r(w2, w3, sh);
      int16_t *hor = reinterpret_cast<int16_t *>(smem_raw + g * slot_bytes + win_bytes + W * h) + (gx * wh + r) * W + c0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t lo = e ? __funnelshift_r(A0, A1, 8 * e) : A0, hi = e ? __funnelshift_r(A1, A2, 8 * e) : A1;
        hor[e] = (int16_t)dp4a_us8(hi, fhi, dp4a_us8(lo, flo, 0));          // 8-bit: shift1 = 0
      }
    }
    __syncthreads();
    // (3) vertical pass + SATD: item = (block, candidate, 8x8 sub-block)
    for (int i = threadIdx.x; i < G * a.k * nsub; i += blockDim.x) {
      const int g = i / (a.k * nsub), rem = i - g * (a.k * nsub), kk = rem / nsub, sb = rem - kk * nsub;
      const int sy = sb / nsx, sx = sb - sy * nsx;
      if (b0 + g >= n_blocks) continue;
      const size_t b = b0 + g;
      const int mvy = a.mvy[b], cy = mvy + a.cdy[kk], rely = (cy >> 4) - (mvy >> 4);
      int fv[8];
      load_filter<8>(cy & 15, fv);
      const int16_t *hor = reinterpret_cast<const int16_t *>(smem_raw + g * slot_bytes + win_bytes + W * h) +
                           (a.cg[kk] * wh + 1 + rely + sy * 8) * W + sx * 8;
      const unsigned char *orig = smem_raw + g * slot_bytes + win_bytes + (sy * 8) * W + sx * 8;
      uint32_t wo[8][2], wc[8][2];
#pragma unroll
      for (int y = 0; y < 8; ++y) {
        int s[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint4 v = *reinterpret_cast<const uint4 *>(hor + (y + t) * W);
          const int x0 = (int)(short)(v.x & 0xffff), x1 = (int)v.x >> 16, x2 = (int)(short)(v.y & 0xffff), x3 = (int)v.y >> 16;
          const int x4 = (int)(short)(v.z & 0xffff), x5 = (int)v.z >> 16, x6 = (int)(short)(v.w & 0xffff), x7 = (int)v.w >> 16;
          s[0] += fv[t] * x0; s[1] += fv[t] * x1; s[2] += fv[t] * x2; s[3] += fv[t] * x3;
          s[4] += fv[t] * x4; s[5] += fv[t] * x5; s[6] += fv[t] * x6; s[7] += fv[t] * x7;
        }
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          const uint32_t p = (uint32_t)min(max(((s[x] >> 6) + 32) >> 6, 0), 255);
          if (x < 4) lo |= p << (8 * x); else hi |= p << (8 * (x - 4));
        }
        wc[y][0] = lo; wc[y][1] = hi;
        const uint2 o = *reinterpret_cast<const uint2 *>(orig + y * W);
        wo[y][0] = o.x; wo[y][1] = o.y;
      }
      unsigned dummy;
      const unsigned part = satd8_words_u8<false>(wo, wc, dummy);
      if (nsub == 1) a.out[b * a.k + kk] = part;
      else atomicAdd(&a.out[b * a.k + kk], part);
    }
  }
}

}  // namespace uvgb

using namespace uvgb;

static int ipol_launch(int bitdepth, int chroma, int hi, const void *plane, int stride, int plane_w, int plane_h, int w,
                       int h, const int32_t *bx, const int32_t *by, const int32_t *mvx, const int32_t *mvy, void *dst,
                       size_t n, cudaStream_t st)
{
  if (bitdepth < 8 || bitdepth > 14) { set_error("ipol: bitdepth %d out of range", bitdepth); return 1; }
  if (w < 1 || h < 1 || w > 64 || h > 64) { set_error("ipol: block %dx%d unsupported", w, h); return 1; }
  if (n == 0) return 0;
  IpolArgs a{ plane, stride, plane_w, plane_h, w, h, bx, by, mvx, mvy, dst, hi, bitdepth - 8, 14 - bitdepth,
              (1 << bitdepth) - 1 };
  const int pad = chroma ? 3 : 7, pix = bitdepth == 8 ? 1 : 2;
  const size_t smem = (((size_t)(w + pad) * (h + pad) * pix + 15) & ~(size_t)15) + (size_t)(h + pad) * w * 2;
  size_t grid = n;
  const size_t cap = (size_t)sm_count() * 16;
  if (grid > cap) grid = cap;
  if (bitdepth == 8) {
    if (chroma) ipol_kernel<uint8_t, 4><<<(unsigned)grid, 128, smem, st>>>(a, n);
    else ipol_kernel<uint8_t, 8><<<(unsigned)grid, 128, smem, st>>>(a, n);
  } else {
    if (chroma) ipol_kernel<uint16_t, 4><<<(unsigned)grid, 128, smem, st>>>(a, n);
    else ipol_kernel<uint16_t, 8><<<(unsigned)grid, 128, smem, st>>>(a, n);
  }
  UVGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int uvg_b200_ipol_dev(int bitdepth, int chroma, int hi, const void *plane, int stride, int plane_w,
                                 int plane_h, int w, int h, const int32_t *bx, const int32_t *by, const int32_t *mvx,
                                 const int32_t *mvy, void *dst, size_t n_blocks, void *stream)
{
  if (ensure_init()) return 1;
  return ipol_launch(bitdepth, chroma, hi, plane, stride, plane_w, plane_h, w, h, bx, by, mvx, mvy, dst, n_blocks,
                     static_cast<cudaStream_t>(stream));
}

extern "C" int uvg_b200_frac_search_dev(int bitdepth, int w, int h, const void *cur, const void *ref, int stride,
                                        int plane_w, int plane_h, const int32_t *bx, const int32_t *by,
                                        const int32_t *mvx, const int32_t *mvy, const int8_t *offsets, int k,
                                        uint32_t *out, size_t n_blocks, void *stream)
{
  if (ensure_init()) return 1;
  if (bitdepth != 8) { set_error("frac_search: 8-bit only; use ipol_dev + distortion_dev"); return 1; }
  if ((w != 8 && w != 16 && w != 32 && w != 64) || h % 8 || h < 8 || h > 64) { set_error("frac_search: block %dx%d unsupported", w, h); return 1; }
  if (k < 1 || k > 64) { set_error("frac_search: 1..64 candidates (got %d)", k); return 1; }
  if (n_blocks == 0) return 0;
  FracArgs a;
  a.cur = static_cast<const uint8_t *>(cur); a.ref = static_cast<const uint8_t *>(ref);
  a.stride = stride; a.plane_w = plane_w; a.plane_h = plane_h; a.h = h;
  a.bx = bx; a.by = by; a.mvx = mvx; a.mvy = mvy; a.k = k; a.out = out; a.nx = 0;
  for (int i = 0; i < k; ++i) {
    const int dx = offsets[2 * i], dy = offsets[2 * i + 1];
    if (dx < -16 || dx > 15 || dy < -16 || dy > 15) { set_error("frac_search: offset out of -16..15"); return 1; }
    int g = 0;
    while (g < a.nx && a.gdx[g] != dx) ++g;
    if (g == a.nx) {
      if (a.nx == 8) { set_error("frac_search: more than 8 distinct horizontal offsets"); return 1; }
      a.gdx[a.nx++] = (signed char)dx;
    }
    a.cg[i] = (signed char)g; a.cdy[i] = (signed char)dy;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nsub = (w / 8) * (h / 8);
  if (nsub > 1) UVGB_TRY(cudaMemsetAsync(out, 0, n_blocks * (size_t)k * sizeof(uint32_t), st));   // sub-blocks accumulate
  const int G = w == 8 ? 8 : w == 16 ? 2 : 1;
  const int P = (w + 10 + 3 + 3) & ~3, wh = h + 10;
  const size_t slot = (size_t)((wh * P + 15) & ~15) + (size_t)w * h + (size_t)((a.nx * wh * w * 2 + 15) & ~15);
  const size_t smem = slot * G;
  if (smem > 200 * 1024) { set_error("frac_search: candidate set needs %zu bytes of shared memory", smem); return 1; }
  // callable from any number of threads: set the attributes exactly once
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    cudaError_t e = cudaFuncSetAttribute(frac_search_kernel<8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(frac_search_kernel<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(frac_search_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(frac_search_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_err = e;
  });
  UVGB_TRY(attr_err);
  size_t grid = (n_blocks + G - 1) / G;
  const size_t cap = (size_t)sm_count() * 16;
  if (grid > cap) grid = cap;
  if (w == 8) frac_search_kernel<8, 8><<<(unsigned)grid, 128, smem, st>>>(a, n_blocks);
  else if (w == 16) frac_search_kernel<16, 2><<<(unsigned)grid, 128, smem, st>>>(a, n_blocks);
  else if (w == 32) frac_search_kernel<32, 1><<<(unsigned)grid, 128, smem, st>>>(a, n_blocks);
  else frac_search_kernel<64, 1><<<(unsigned)grid, 128, smem, st>>>(a, n_blocks);
  UVGB_LAUNCH_CHECK();
  return 0;
}

// ---- drop-in leaves: strategies-ipol.h:105-108 (host pointers, synchronous) -----------------------------
namespace {
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <typename Pix>
void sample_leaf(int chroma, int hi, const Pix *src, int src_stride, int w, int h, void *dst, int dst_stride, int mvx, int mvy)
{
  if (ensure_init()) { fprintf(stderr, "uvg266_b200: %s\n", uvg_b200_last_error()); abort(); }
  if (w < 1 || h < 1 || w > 64 || h > 64) { fprintf(stderr, "uvg266_b200: sample interpolation: block %dx%d unsupported\n", w, h); abort(); }
  const int off = chroma ? 1 : 3, pad = chroma ? 3 : 7, frac = chroma ? 5 : 4;
  const int ww = w + pad, wh = h + pad;
  Staging &s = tls_staging();
  const size_t win_bytes = align_up((size_t)ww * wh * sizeof(Pix), 256), out_elem = hi ? 2 : sizeof(Pix);
  const size_t out_bytes = align_up((size_t)w * h * out_elem, 256);
  if (s.ensure(win_bytes + out_bytes + 64, win_bytes + out_bytes + 256)) { fprintf(stderr, "uvg266_b200: %s\n", uvg_b200_last_error()); abort(); }
  // pack the window rows (the caller's buffer is strided), one copy in, one kernel, one copy out
  Pix *hw = reinterpret_cast<Pix *>(s.h_pin);
  for (int r = 0; r < wh; ++r) memcpy(hw + (size_t)r * ww, src + (ptrdiff_t)(r - off) * src_stride - off, (size_t)ww * sizeof(Pix));
  UVGB_MUST(cudaMemcpyAsync(s.d_buf, hw, (size_t)ww * wh * sizeof(Pix), cudaMemcpyHostToDevice, s.stream));
  int32_t *d_job = reinterpret_cast<int32_t *>(s.d_buf + win_bytes + out_bytes);
  const int32_t job[4] = { off, off, mvx & ((1 << frac) - 1), mvy & ((1 << frac) - 1) };
  int32_t *h_job = reinterpret_cast<int32_t *>(s.h_pin + win_bytes + out_bytes);
  memcpy(h_job, job, sizeof(job));
  UVGB_MUST(cudaMemcpyAsync(d_job, h_job, sizeof(job), cudaMemcpyHostToDevice, s.stream));
  const int bd = dropin_bitdepth(sizeof(Pix));
  if (ipol_launch(bd, chroma, hi, s.d_buf, ww, ww, wh, w, h, d_job, d_job + 1, d_job + 2, d_job + 3, s.d_buf + win_bytes, 1, s.stream)) {
    fprintf(stderr, "uvg266_b200: %s\n", uvg_b200_last_error()); abort();
  }
  uint8_t *ho = s.h_pin + win_bytes;
  UVGB_MUST(cudaMemcpyAsync(ho, s.d_buf + win_bytes, (size_t)w * h * out_elem, cudaMemcpyDeviceToHost, s.stream));
  UVGB_MUST(cudaStreamSynchronize(s.stream));
  for (int r = 0; r < h; ++r)
    memcpy(static_cast<uint8_t *>(dst) + (size_t)r * dst_stride * out_elem, ho + (size_t)r * w * out_elem, (size_t)w * out_elem);
}
}  // namespace

// The four FME block filters (strategies-ipol.h:142-145, ipol-generic.c:213-680): each of their four output
// blocks is the sample interpolation of the block at one quarter-pel offset from its integer position
// (sample (1,1) of src) - established against the reference by tests/test_oracle_ipol_vs_ref.py:
//   step 0: (-2,0) (2,0) (0,-2) (0,2)          step 1: (-2,-2) (2,-2) (-2,2) (2,2)
//   step 2: c+(-1,0) c+(1,0) c+(0,-1) c+(0,1)   step 3: c+(-1,-1) c+(1,-1) c+(-1,1) c+(1,1),  c = 2 * (off_x, off_y)
// so one launch of ipol_kernel with four jobs over the same window serves a call; the reference's
// intermediate buffers are not needed (and not written: install the four filters together).
namespace {
template <typename Pix>
void fme_leaf(int step, const Pix *src, int src_stride, int w, int h, Pix *filtered, int filtered_stride, int block_pitch,
              int off_x, int off_y)
{
  static const signed char d[4][4][2] = { { { -2, 0 }, { 2, 0 }, { 0, -2 }, { 0, 2 } }, { { -2, -2 }, { 2, -2 }, { -2, 2 }, { 2, 2 } },
                                          { { -1, 0 }, { 1, 0 }, { 0, -1 }, { 0, 1 } }, { { -1, -1 }, { 1, -1 }, { -1, 1 }, { 1, 1 } } };
  if (ensure_init()) { fprintf(stderr, "uvg266_b200: %s\n", uvg_b200_last_error()); abort(); }
  if (step < 0 || step > 3 || w < 1 || h < 1 || w > 64 || h > 64) { fprintf(stderr, "uvg266_b200: fme filter: bad step / block %dx%d\n", w, h); abort(); }
  const int ww = w + 8, wh = h + 8;                        // samples -3 .. w+4 around the (w+1) x (h+1) area
  Staging &s = tls_staging();
  const size_t win_bytes = align_up((size_t)ww * wh * sizeof(Pix), 256), out_bytes = align_up((size_t)4 * w * h * sizeof(Pix), 256);
  if (s.ensure(win_bytes + out_bytes + 64, win_bytes + out_bytes + 256)) { fprintf(stderr, "uvg266_b200: %s\n", uvg_b200_last_error()); abort(); }
  Pix *hw = reinterpret_cast<Pix *>(s.h_pin);
  for (int r = 0; r < wh; ++r) memcpy(hw + (size_t)r * ww, src + (ptrdiff_t)(r - 3) * src_stride - 3, (size_t)ww * sizeof(Pix));
  int32_t *h_job = reinterpret_cast<int32_t *>(s.h_pin + win_bytes + out_bytes);
  for (int j = 0; j < 4; ++j) {
    const int qx = (step >= 2 ? 2 * off_x : 0) + d[step][j][0], qy = (step >= 2 ? 2 * off_y : 0) + d[step][j][1];
    h_job[j] = 4; h_job[4 + j] = 4;                       // the integer position (1,1) of src inside the window
    h_job[8 + j] = qx * 4; h_job[12 + j] = qy * 4;        // quarter-pel -> 1/16 samples
  }
  UVGB_MUST(cudaMemcpyAsync(s.d_buf, hw, (size_t)ww * wh * sizeof(Pix), cudaMemcpyHostToDevice, s.stream));
  int32_t *d_job = reinterpret_cast<int32_t *>(s.d_buf + win_bytes + out_bytes);
  UVGB_MUST(cudaMemcpyAsync(d_job, h_job, 16 * sizeof(int32_t), cudaMemcpyHostToDevice, s.stream));
  if (ipol_launch(dropin_bitdepth(sizeof(Pix)), 0, 0, s.d_buf, ww, ww, wh, w, h, d_job, d_job + 4, d_job + 8, d_job + 12,
                  s.d_buf + win_bytes, 4, s.stream)) {
    fprintf(stderr, "uvg266_b200: %s\n", uvg_b200_last_error()); abort();
  }
  Pix *ho = reinterpret_cast<Pix *>(s.h_pin + win_bytes);
  UVGB_MUST(cudaMemcpyAsync(ho, s.d_buf + win_bytes, (size_t)4 * w * h * sizeof(Pix), cudaMemcpyDeviceToHost, s.stream));
  UVGB_MUST(cudaStreamSynchronize(s.stream));
  for (int j = 0; j < 4; ++j)
    for (int r = 0; r < h; ++r)
      memcpy(filtered + (size_t)j * block_pitch + (size_t)r * filtered_stride, ho + ((size_t)j * h + r) * w, (size_t)w * sizeof(Pix));
}
}  // namespace

#define UVGB_IPOL_LEAVES(SUF, PIX)                                                                                   \
  extern "C" void uvg_b200_sample_quarterpel_luma_##SUF(const PIX *src, int src_stride, int w, int h, PIX *dst,       \
                                                        int dst_stride, int mvx, int mvy)                           \
  { sample_leaf<PIX>(0, 0, src, src_stride, w, h, dst, dst_stride, mvx, mvy); }                                      \
  extern "C" void uvg_b200_sample_quarterpel_luma_hi_##SUF(const PIX *src, int src_stride, int w, int h, int16_t *dst, \
                                                           int dst_stride, int mvx, int mvy)                        \
  { sample_leaf<PIX>(0, 1, src, src_stride, w, h, dst, dst_stride, mvx, mvy); }                                      \
  extern "C" void uvg_b200_sample_octpel_chroma_##SUF(const PIX *src, int src_stride, int w, int h, PIX *dst,         \
                                                      int dst_stride, int mvx, int mvy)                             \
  { sample_leaf<PIX>(1, 0, src, src_stride, w, h, dst, dst_stride, mvx, mvy); }                                      \
  extern "C" void uvg_b200_sample_octpel_chroma_hi_##SUF(const PIX *src, int src_stride, int w, int h, int16_t *dst,  \
                                                         int dst_stride, int mvx, int mvy)                          \
  { sample_leaf<PIX>(1, 1, src, src_stride, w, h, dst, dst_stride, mvx, mvy); }
UVGB_IPOL_LEAVES(8, uint8_t)
UVGB_IPOL_LEAVES(16, uint16_t)

extern "C" void uvg_b200_fme_filter_8(int step, const uint8_t *src, int src_stride, int w, int h, uint8_t *filtered,
                                      int filtered_stride, int block_pitch, int off_x, int off_y)
{ fme_leaf<uint8_t>(step, src, src_stride, w, h, filtered, filtered_stride, block_pitch, off_x, off_y); }
extern "C" void uvg_b200_fme_filter_16(int step, const uint16_t *src, int src_stride, int w, int h, uint16_t *filtered,
                                       int filtered_stride, int block_pitch, int off_x, int off_y)
{ fme_leaf<uint16_t>(step, src, src_stride, w, h, filtered, filtered_stride, block_pitch, off_x, off_y); }
