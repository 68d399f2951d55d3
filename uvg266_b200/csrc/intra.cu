// intra.cu - luma intra prediction (planar / DC / 65 angular modes incl. wide angles, reference
// smoothing, PDPC) on the device, and the fused "predict + score" rough-search kernel.
// This is the first "next" row of the scope table: the producer of the
// predictions the rough search scores with SATD/SAD, and the reference encoder's #1 CPU hotspot.
//
// Restates, for multi_ref_idx = 0, no ISP, no MIP, intra_smoothing enabled:
//   intra_predict_regular            intra.c:660-753   (smoothing decision, dispatch, PDPC)
//   uvg_wide_angle_correction        intra.c:637-658
//   intra_filter_reference           intra.c:190-225   ([1 2 1]/4 reference smoothing)
//   intra_pred_dc                    intra.c:236-273
//   uvg_angular_pred_generic         intra-generic.c:55-299
//   uvg_intra_pred_planar_generic    intra-generic.c:306-360
//   uvg_pdpc_planar_dc_generic       intra-generic.c:414-437
// Every output pixel is a pure function of the two reference rows, so a pixel (or an 8x8
// sub-block) is one thread's work and the prediction never has to exist in HBM: the fused kernel
// feeds predicted rows straight into the packed Hadamard of hadamard.cuh.
//
// Reference rows per block (`refs`): top[2n+1] then left[2n+1], index 0 = the top-left corner
// sample, each row stored with a pitch of 2n+4 pixels; like the encoder's own buffers the rows
// continue past 2n with the last sample replicated (intra.c:879-886).
#include "common.cuh"
#include "hadamard.cuh"
#include "intra.cuh"

namespace uvgb {

// CTA-cooperative: stage the reference rows of block `b` (global layout: top then left, pitch
// 2n+4) into shared memory with replication padding, build the smoothed rows and the DC value.
template <typename Pix>
__device__ __forceinline__ void stage_refs(const Pix *__restrict__ refs, size_t b, int n, int spitch, Pix *s_rows,
                                           int *s_dc, int tid, int nthreads)
{
  const int gp = 2 * n + 4, len = 2 * n + 1;
  const Pix *g = refs + b * (size_t)(2 * gp);
  Pix *top = s_rows, *left = s_rows + spitch, *ftop = s_rows + 2 * spitch, *fleft = s_rows + 3 * spitch;
  for (int i = tid; i < 2 * spitch; i += nthreads) {
    const int row = i / spitch, x = i - row * spitch;
    s_rows[i] = g[row * gp + min(x, len - 1)];
  }
  __syncthreads();
  for (int i = tid; i < 2 * spitch; i += nthreads) {         // intra_filter_reference :190-225
    const int row = i / spitch, x = i - row * spitch;
    const Pix *src = row ? left : top;
    int v;
    if (x == 0) v = ((int)left[1] + 2 * (int)left[0] + (int)top[1] + 2) >> 2;
    else if (x < len - 1) v = ((int)src[x - 1] + 2 * (int)src[x] + (int)src[x + 1] + 2) >> 2;
    else v = src[len - 1];
    (row ? fleft : ftop)[x] = (Pix)v;
  }
  if (tid == 0) {                                            // intra_pred_dc :236-273, square block
    int sum = 0;
    for (int i = 0; i < n; ++i) sum += (int)top[i + 1] + (int)left[i + 1];
    *s_dc = (sum + n) >> (ilog2_dev(2 * n));
  }
  __syncthreads();
}

// ---- predictions out (generic: any n in 4..32, any bit depth) -------------------------------------
template <typename Pix>
__global__ void __launch_bounds__(256)
intra_predict_kernel(const Pix *__restrict__ refs, Pix *__restrict__ preds, size_t n_blocks, int n, ModeTable mt,
                     int pixel_max)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int spitch = 2 * n + 12;
  Pix *s_rows = reinterpret_cast<Pix *>(smem_raw);
  __shared__ int s_dc;
  const int l2 = ilog2_dev(n);
  for (size_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
    __syncthreads();
    stage_refs<Pix>(refs, b, n, spitch, s_rows, &s_dc, threadIdx.x, blockDim.x);
    RefRows<Pix> rr{ s_rows, s_rows + spitch, s_rows + 2 * spitch, s_rows + 3 * spitch };
    const int total = mt.k * n * n;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
      const int m = e / (n * n), p = e - m * n * n, Y = p / n, X = p - Y * n;
      preds[(b * mt.k + m) * (size_t)(n * n) + p] = (Pix)pred_pixel<Pix>(mt.m[m], rr, n, l2, X, Y, s_dc, pixel_max);
    }
  }
}

// ---- fused rough search: predict + min(SATD, 2*SAD), 8-bit, n in {8, 16, 32} ----------------------
// The predictions never exist in HBM: each thread generates 8x8 sub-blocks of a prediction row by
// row in registers and feeds them to the packed Hadamard of hadamard.cuh.
//
// Work mapping. A CTA of 4 warps owns BPC = 32 / (N/8) consecutive blocks, staged once in shared
// memory (reference rows + their smoothed copies, the original block and its transpose). Lane =
// (block, 8-pixel column strip sx); the lane walks its strip top to bottom (sy = 0..N/8-1). Warp w
// takes modes w, w+4, ...: a warp is always inside ONE mode and ONE row index, so the mode
// decisions (direction, filter, PDPC), the row's displacement (delta_int, delta_fract) and the
// interpolation filter are warp-uniform - no divergence.
// Horizontal modes are generated in their natural transposed orientation (main = left row) and
// scored against the transposed original: SAD and the 2-D Hadamard abs-sum do not change under
// transposition, so there is no transposing store (intra-generic.c:286-298 flips a temp block).
// One output pixel of the 4-tap interpolation is one dp4a: the row's 11-byte window of the main
// reference is fetched as 4 aligned words and funnel-shifted; the filter is a packed s8x4 word.
// Negative angles extend the main reference with the projected side reference (:145-156) into a
// small per-warp scratch row that ends exactly where index 0 of the main row begins.
template <int N>
struct RoughCfg {
  static constexpr int NS = N / 8;                   // 8-pixel strips per block row
  static constexpr int BPC = 32 / NS;                // blocks per CTA
  static constexpr int RS = 2 * N + 12;              // shared-memory pitch of one reference row
  static constexpr int BS = 4 * RS + 4;              // per-block stride of the 4 rows: odd word count
  static constexpr int HEAD = N + 8;                 // scratch for main-reference indices -HEAD..-1
  static constexpr int RP = N == 8 ? 32 : N == 16 ? 64 : 128;   // RS rounded up to a power of two (index math of the prologue)
  // bytes of one (warp, block) extension scratch: HEAD + mirror, rounded to an ODD multiple of 8 bytes (8-byte stores stay
  // aligned, lanes of different blocks spread over the banks: 32 bytes for 8x8 was an 8-way conflict)
  static constexpr int EXT8 = (HEAD + kExtMirror + 7) / 8;
  static constexpr int EXTROW = 8 * (EXT8 | 1);
  static constexpr int OP = N * N + 8 * NS;          // per-block stride of the original (bank spread)
};

// PLANE: block gb is CTU-major block (gb % per_ctu) of CTU ctu_begin + gb / per_ctu of the picture `src` (frame.cu);
// original and reference rows come from the plane with border replication, `A` / `refs` are unused.
template <int N, bool PLANE>
__global__ void __launch_bounds__(128, N == 8 ? 8 : N == 16 ? 6 : 5)    // 64 / 80 / 96 registers
intra_rough_kernel(Blocks A, const uint8_t *__restrict__ refs, uint32_t *__restrict__ out, size_t n_blocks, ModeTable mt,
                   PlaneSrc src)
{
  using Cfg = RoughCfg<N>;
  constexpr int NS = Cfg::NS, BPC = Cfg::BPC, RS = Cfg::RS, BS = Cfg::BS, HEAD = Cfg::HEAD, OP = Cfg::OP, RP = Cfg::RP;
  constexpr int GP = 2 * N + 4, L2 = N == 8 ? 3 : N == 16 ? 4 : 5;
  __shared__ __align__(16) uint8_t s_rows[BPC * BS];
  __shared__ __align__(16) uint8_t s_ext[4][BPC][Cfg::EXTROW];   // [.. HEAD) projected side samples, then a copy of main[0..12)
  __shared__ __align__(16) uint8_t s_orig[BPC * OP], s_origT[BPC * OP];
  __shared__ int s_dc[BPC];
  __shared__ int s_X[BPC], s_Y[BPC];
  const size_t first = (size_t)blockIdx.x * BPC;
  const int tid = threadIdx.x;
  if (PLANE) {
    if (tid < BPC) {
      constexpr int BPR = 64 / N, PPC = BPR * BPR;
      const size_t gb = min(first + tid, n_blocks - 1);
      const int ctu = (int)(gb / PPC) + src.ctu_begin, j = (int)(gb % PPC);
      s_X[tid] = (ctu % src.ctu_cols) * 64 + (j % BPR) * N;
      s_Y[tid] = (ctu / src.ctu_cols) * 64 + (j / BPR) * N;
    }
    __syncthreads();
  }
  // reference rows (tail replicated)
  for (int i = tid; i < BPC * 2 * RP; i += 128) {                        // (power-of-two index space: shifts, no divisions)
    const int blk = i / (2 * RP), row = (i / RP) & 1, x = i & (RP - 1);
    if (x >= RS) continue;
    if (PLANE) {
      const int xx = min(x, 2 * N), X = s_X[blk], Y = s_Y[blk];
      const int sx = row ? X - 1 : X - 1 + xx, sy = row ? Y - 1 + xx : Y - 1;
      s_rows[blk * BS + row * RS + x] = __ldg(src.plane + (size_t)min(max(sy, 0), src.H - 1) * src.stride + min(max(sx, 0), src.W - 1));
    } else {
      const size_t gb = min(first + blk, n_blocks - 1);
      s_rows[blk * BS + row * RS + x] = __ldg(refs + gb * (size_t)(2 * GP) + row * GP + min(x, 2 * N));
    }
  }
  // original blocks, as they are and transposed
  for (int i = tid; i < BPC * N * N; i += 128) {
    const int blk = i / (N * N), p = i - blk * N * N, y = p / N, x = p - y * N;
    uint8_t v;
    if (PLANE) {
      v = __ldg(src.plane + (size_t)min(s_Y[blk] + y, src.H - 1) * src.stride + min(s_X[blk] + x, src.W - 1));
    } else {
      const size_t gb = min(first + blk, n_blocks - 1);
      v = __ldg(static_cast<const uint8_t *>(A.base) + A.offset(gb) + (size_t)y * A.stride + x);
    }
    s_orig[blk * OP + p] = v;
    s_origT[blk * OP + x * N + y] = v;
  }
  __syncthreads();
  for (int i = tid; i < BPC * 2 * RP; i += 128) {                         // intra_filter_reference :190-225
    const int blk = i / (2 * RP), row = (i / RP) & 1, x = i & (RP - 1);
    if (x >= RS) continue;
    const uint8_t *top = s_rows + blk * BS, *left = top + RS, *src = row ? left : top;
    int v;
    if (x == 0) v = ((int)left[1] + 2 * (int)left[0] + (int)top[1] + 2) >> 2;
    else if (x < 2 * N) v = ((int)src[x - 1] + 2 * (int)src[x] + (int)src[x + 1] + 2) >> 2;
    else v = src[2 * N];
    s_rows[blk * BS + (2 + row) * RS + x] = (uint8_t)v;
  }
  if (tid < BPC) {                                                         // intra_pred_dc :236-273
    const uint8_t *top = s_rows + tid * BS, *left = top + RS;
    int sum = 0;
    for (int i = 1; i <= N; ++i) sum += (int)top[i] + (int)left[i];
    s_dc[tid] = (sum + N) >> (L2 + 1);
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31, blk = lane / NS, sx = lane % NS;
  const size_t gb = first + blk;
  const uint8_t *rows = s_rows + blk * BS;
  uint8_t *ext0 = &s_ext[warp][blk][0] + HEAD;
  for (int m = warp; m < mt.k; m += 4) {
    const ModeParam mp = mt.m[m];
    const uint8_t *top = rows + (mp.filtered ? 2 * RS : 0), *left = top + RS;
    unsigned satd = 0, sad = 0;
    if (mp.kind == 2) {
      const uint8_t *mainr = mp.vertical ? top : left, *side = mp.vertical ? left : top;
      const uint8_t *orig = (mp.vertical ? s_orig : s_origT) + blk * OP + 8 * sx;
      if (mp.sample_disp < 0) {
        __syncwarp();                                                     // previous mode's reads of the scratch
        uint32_t e[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t w = 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = 8 * sx + 8 - (4 * h + q);                      // main index -j
            w |= (uint32_t)side[min((j * mp.inv_angle + 256) >> 9, N)] << (8 * q);
          }
          e[h] = w;
        }
        *reinterpret_cast<uint2 *>(ext0 - 8 * sx - 8) = make_uint2(e[0], e[1]);
        if (sx == 0) {                                                    // angular_row8 reads a row's words from one base
#pragma unroll
          for (int q = 0; q < kExtMirror / 4; ++q) reinterpret_cast<uint32_t *>(ext0)[q] = reinterpret_cast<const uint32_t *>(mainr)[q];
        }
        __syncwarp();
      }
      const int lim = min(3 << mp.scale, N);
      const PdpcTab tab = mp.pdpc ? pdpc_tab(mp, sx, lim) : PdpcTab{};
#pragma unroll 1
      for (int sy = 0; sy < NS; ++sy) {
        uint32_t wo[8][2], wc[8][2];
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int r = sy * 8 + rr;
          const uint2 o = *reinterpret_cast<const uint2 *>(orig + r * N);
          wo[rr][0] = o.x; wo[rr][1] = o.y;
          angular_row8<true>(mp, mainr, side, ext0, r, sx, lim, wc[rr][0], wc[rr][1], tab);
        }
        unsigned s1 = 0;
        satd += satd8_words_u8<true>(wo, wc, s1);
        sad += s1;
      }
    } else {
      const uint8_t *orig = s_orig + blk * OP + 8 * sx;
      const int dc = s_dc[blk], tr = top[N + 1], bl = left[N + 1];
#pragma unroll 1
      for (int sy = 0; sy < NS; ++sy) {
        uint32_t wo[8][2], wc[8][2];
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int Y = sy * 8 + rr, lY = left[Y + 1];
          const uint2 o = *reinterpret_cast<const uint2 *>(orig + Y * N);
          wo[rr][0] = o.x; wo[rr][1] = o.y;
          const int wT = 32 >> min(31, (Y << 1) >> mp.scale);
          uint32_t w2[2] = { 0, 0 };
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int X = 8 * sx + c, tX = top[X + 1];
            int v = dc;
            if (mp.kind == 0) {                                           // planar :306-360
              const int hor = (lY << L2) + (X + 1) * (tr - lY), ver = (tX << L2) + (Y + 1) * (bl - tX);
              v = ((hor << L2) + (ver << L2) + (1 << (2 * L2))) >> (1 + 2 * L2);
            }
            const int wL = 32 >> min(31, (X << 1) >> mp.scale);           // PDPC :414-437 (n >= 4 always)
            v = (v + ((wL * (lY - v) + wT * (tX - v) + 32) >> 6)) & 255;
            w2[c >> 2] |= (uint32_t)v << (8 * (c & 3));
          }
          wc[rr][0] = w2[0]; wc[rr][1] = w2[1];
        }
        unsigned s1 = 0;
        satd += satd8_words_u8<true>(wo, wc, s1);
        sad += s1;
      }
    }
    if (NS > 1) { satd = group_sum<NS>(satd); sad = group_sum<NS>(sad); }
    if (sx == 0 && gb < n_blocks) out[gb * mt.k + m] = min(satd, 2 * sad);
  }
}

// ---- strategy leaves with all their arguments (drop-ins, one block per call) ------------------------
// uvg_angular_pred (intra-generic.c:55-299), uvg_intra_pred_planar (:306-360), uvg_pdpc_planar_dc
// (:414-437) exactly as the encoder calls them from intra_predict_regular: any w x h in 4..64 (2..64
// for chroma), wide-angle corrected modes, chroma's 2-tap filter, multi_ref_idx, isp_mode.
struct LeafParams { int kind, w, h, mode, channel, mrl, isp, pixel_max, ref_len; };   // kind 0 angular, 1 planar, 2 pdpc

template <typename Pix>
__global__ void __launch_bounds__(256)
intra_leaf_kernel(LeafParams p, const Pix *__restrict__ above_g, const Pix *__restrict__ left_g, Pix *__restrict__ dst)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Pix *above = reinterpret_cast<Pix *>(smem_raw), *left = above + p.ref_len;
  for (int i = threadIdx.x; i < p.ref_len; i += blockDim.x) { above[i] = above_g[i]; left[i] = left_g[i]; }
  __syncthreads();
  const int w = p.w, h = p.h, l2w = ilog2_dev(w), l2h = ilog2_dev(h);
  const int pmask = sizeof(Pix) == 1 ? 0xff : 0xffff;
  if (p.kind == 1) {
    const int tr = above[w + 1], bl = left[h + 1];
    for (int e = threadIdx.x; e < w * h; e += blockDim.x) {
      const int y = e / w, x = e - y * w;
      const int hor = ((int)left[y + 1] << l2w) + (x + 1) * (tr - (int)left[y + 1]);
      const int ver = ((int)above[x + 1] << l2h) + (y + 1) * (bl - (int)above[x + 1]);
      dst[e] = (Pix)(((hor << l2h) + (ver << l2w) + (1 << (l2w + l2h))) >> (1 + l2w + l2h));
    }
    return;
  }
  if (p.kind == 2) {
    const int scale = (l2w + l2h - 2) >> 2;
    for (int e = threadIdx.x; e < w * h; e += blockDim.x) {
      const int y = e / w, x = e - y * w;
      const int wT = 32 >> min(31, (y << 1) >> scale), wL = 32 >> min(31, (x << 1) >> scale);
      const int cur = dst[e];
      dst[e] = (Pix)((cur + ((wL * ((int)left[y + 1] - cur) + wT * ((int)above[x + 1] - cur) + 32) >> 6)) & pmask);
    }
    return;
  }
  const int mode = p.mode, mrl = p.mrl;
  const bool vertical = mode >= 34;
  const int mode_disp = vertical ? mode - 50 : 18 - mode, ad = abs(mode_disp);
  const int sd = (mode_disp < 0 ? -1 : 1) * c_sampledisp[ad], inv = c_invsampledisp[ad];
  const int scale = min(2, (vertical ? l2h : l2w) - c_pre_scale[ad]);
  const int W = vertical ? w : h, H = vertical ? h : w;              // work block: horizontal modes are transposed
  const Pix *main_in = vertical ? above : left, *side_in = vertical ? left : above;
  auto MAINP = [&](int j) -> int {                                    // ref_main[j] after "ref_main += mrl"
    const int idx = j + mrl;
    return idx >= 0 ? main_in[idx] : side_in[min((-idx * inv + 256) >> 9, H)];
  };
  const bool frac = (abs(sd) & 31) != 0;
  bool cubic = true;
  if (min(abs(mode - 50), abs(mode - 18)) > c_dist_thres[(l2w + l2h) >> 1] && frac) cubic = false;
  if (mrl || p.isp) cubic = true;
  bool pdpc = W >= 4 && H >= 4 && mrl == 0;
  if (sd != 0) {
    if (mode > 1 && mode < 67) {
      if (mode_disp < 0) pdpc = false;
      else if (mode_disp > 0) pdpc = pdpc && scale >= 0;
    }
    pdpc = pdpc && scale >= 0;                                        // (3 << negative) is undefined in the reference
  }
  const int sc0 = (l2w + l2h - 2) >> 2;
  for (int e = threadIdx.x; e < W * H; e += blockDim.x) {
    const int r = e / W, c = e - r * W;
    int v;
    if (sd == 0) {
      v = MAINP(1 + c);
      if (pdpc && c < min(3 << sc0, W))
        v = clip3(0, p.pixel_max, v + (((32 >> ((2 * c) >> sc0)) * ((int)side_in[1 + r + mrl] - MAINP(0)) + 32) >> 6));
    } else {
      const int dpos = (r + 1 + mrl) * sd, di = dpos >> 5, fr = dpos & 31;
      if (!frac) {
        v = MAINP(c + di + 1);
      } else if (p.channel == 0) {
        int f0, f1, f2, f3;
        if (cubic) { f0 = c_cubic[fr][0]; f1 = c_cubic[fr][1]; f2 = c_cubic[fr][2]; f3 = c_cubic[fr][3]; }
        else { f0 = 16 - (fr >> 1); f1 = 32 - (fr >> 1); f2 = 16 + (fr >> 1); f3 = fr >> 1; }
        v = clip3(0, p.pixel_max, (f0 * MAINP(di + c) + f1 * MAINP(di + c + 1) + f2 * MAINP(di + c + 2) + f3 * MAINP(di + c + 3) + 32) >> 6);
      } else {
        const int r1 = MAINP(c + di + 1), r2 = MAINP(c + di + 2);
        v = (r1 + ((fr * (r2 - r1) + 16) >> 5)) & pmask;
      }
      if (pdpc && c < min(3 << scale, W)) {
        const int l = side_in[r + ((256 + (c + 1) * inv) >> 9) + 1 + mrl];
        v = (v + (((32 >> ((2 * c) >> scale)) * (l - v) + 32) >> 6)) & pmask;
      }
    }
    dst[vertical ? r * w + c : c * w + r] = (Pix)v;
  }
}

static int fill_modes(ModeTable &mt, int n, const int8_t *modes, int k)
{
  if (n != 4 && n != 8 && n != 16 && n != 32) { set_error("intra: block size %d unsupported", n); return 1; }
  if (k < 1 || k > kMaxModes) { set_error("intra: 1..67 modes per call (got %d)", k); return 1; }
  mt.k = k;
  for (int i = 0; i < k; ++i) {
    if (modes[i] < 0 || modes[i] > 66) { set_error("intra: mode %d out of range", modes[i]); return 1; }
    mt.m[i] = make_mode(n, modes[i]);
  }
  return 0;
}

}  // namespace uvgb

using namespace uvgb;

extern "C" int uvg_b200_intra_predict_dev(int bitdepth, int n, const void *refs, const int8_t *modes, int k,
                                          void *preds, size_t n_blocks, void *stream)
{
  if (ensure_init()) return 1;
  if (bitdepth < 8 || bitdepth > 16) { set_error("bitdepth %d out of range", bitdepth); return 1; }
  ModeTable mt;
  if (fill_modes(mt, n, modes, k)) return 1;
  if (n_blocks == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int pix = bitdepth == 8 ? 1 : 2;
  const size_t smem = (size_t)4 * (2 * n + 12) * pix;
  size_t grid = n_blocks;
  const size_t cap = (size_t)sm_count() * 16;
  if (grid > cap) grid = cap;
  if (bitdepth == 8)
    intra_predict_kernel<uint8_t><<<(unsigned)grid, 256, smem, st>>>(static_cast<const uint8_t *>(refs),
                                                                     static_cast<uint8_t *>(preds), n_blocks, n, mt, 255);
  else
    intra_predict_kernel<uint16_t><<<(unsigned)grid, 256, smem, st>>>(static_cast<const uint16_t *>(refs),
                                                                      static_cast<uint16_t *>(preds), n_blocks, n, mt,
                                                                      (1 << bitdepth) - 1);
  UVGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int uvg_b200_intra_rough_dev(int bitdepth, int n, const uvg_b200_blocks *orig, const void *refs,
                                        const int8_t *modes, int k, uint32_t *out, size_t n_blocks, void *stream)
{
  if (ensure_init()) return 1;
  if (bitdepth != 8) { set_error("intra_rough: fused kernel is 8-bit; use intra_predict_dev + rough_cost_dev"); return 1; }
  if (n != 8 && n != 16 && n != 32) { set_error("intra_rough: n must be 8, 16 or 32"); return 1; }
  ModeTable mt;
  if (fill_modes(mt, n, modes, k)) return 1;
  if (n_blocks == 0) return 0;
  const Blocks A = to_blocks(orig);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint8_t *r = static_cast<const uint8_t *>(refs);
  auto grid = [&](int bpc) { return (unsigned)((n_blocks + bpc - 1) / bpc); };
  const PlaneSrc none{};
  if (n == 8) intra_rough_kernel<8, false><<<grid(RoughCfg<8>::BPC), 128, 0, st>>>(A, r, out, n_blocks, mt, none);
  else if (n == 16) intra_rough_kernel<16, false><<<grid(RoughCfg<16>::BPC), 128, 0, st>>>(A, r, out, n_blocks, mt, none);
  else intra_rough_kernel<32, false><<<grid(RoughCfg<32>::BPC), 128, 0, st>>>(A, r, out, n_blocks, mt, none);
  UVGB_LAUNCH_CHECK();
  return 0;
}

int uvgb::launch_intra_rough_plane(int n, const PlaneSrc &src, const int8_t *modes, int k, uint32_t *out, size_t n_blocks,
                                   cudaStream_t st)
{
  if (n != 8 && n != 16 && n != 32) { set_error("intra_rough: n must be 8, 16 or 32"); return 1; }
  ModeTable mt;
  if (fill_modes(mt, n, modes, k)) return 1;
  if (n_blocks == 0) return 0;
  const Blocks none{};
  auto grid = [&](int bpc) { return (unsigned)((n_blocks + bpc - 1) / bpc); };
  if (n == 8) intra_rough_kernel<8, true><<<grid(RoughCfg<8>::BPC), 128, 0, st>>>(none, nullptr, out, n_blocks, mt, src);
  else if (n == 16) intra_rough_kernel<16, true><<<grid(RoughCfg<16>::BPC), 128, 0, st>>>(none, nullptr, out, n_blocks, mt, src);
  else intra_rough_kernel<32, true><<<grid(RoughCfg<32>::BPC), 128, 0, st>>>(none, nullptr, out, n_blocks, mt, src);
  UVGB_LAUNCH_CHECK();
  return 0;
}

// ---- drop-in leaves (host pointers, synchronous, per-thread stream + staging) -----------------------
namespace {
constexpr int kRefLen = 2 * 128 + 3 + 33 * 3;      // uvg_intra_ref rows (intra.h:47-50)
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <typename Pix>
void leaf_call(const LeafParams &lp, const void *above, const void *left, void *dst, bool dst_in)
{
  if (ensure_init()) { fprintf(stderr, "uvg266_b200: %s\n", uvg_b200_last_error()); abort(); }
  Staging &s = tls_staging();
  const size_t rb = align_up(kRefLen * sizeof(Pix), 256), db = (size_t)lp.w * lp.h * sizeof(Pix);
  if (s.ensure(0, 2 * rb + align_up(db, 256))) { fprintf(stderr, "uvg266_b200: %s\n", uvg_b200_last_error()); abort(); }
  UVGB_MUST(cudaMemcpyAsync(s.d_buf, above, kRefLen * sizeof(Pix), cudaMemcpyHostToDevice, s.stream));
  UVGB_MUST(cudaMemcpyAsync(s.d_buf + rb, left, kRefLen * sizeof(Pix), cudaMemcpyHostToDevice, s.stream));
  Pix *d_dst = reinterpret_cast<Pix *>(s.d_buf + 2 * rb);
  if (dst_in) UVGB_MUST(cudaMemcpyAsync(d_dst, dst, db, cudaMemcpyHostToDevice, s.stream));
  intra_leaf_kernel<Pix><<<1, 256, 2 * kRefLen * sizeof(Pix), s.stream>>>(lp, reinterpret_cast<const Pix *>(s.d_buf),
                                                                        reinterpret_cast<const Pix *>(s.d_buf + rb), d_dst);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  UVGB_MUST(cudaGetLastError());
  UVGB_MUST(cudaMemcpyAsync(dst, d_dst, db, cudaMemcpyDeviceToHost, s.stream));
  UVGB_MUST(cudaStreamSynchronize(s.stream));
}
void leaf_check(int w, int h, const char *what)
{
  const bool okw = w == 2 || w == 4 || w == 8 || w == 16 || w == 32 || w == 64;
  const bool okh = h == 2 || h == 4 || h == 8 || h == 16 || h == 32 || h == 64;
  if (!okw || !okh) { fprintf(stderr, "uvg266_b200: %s: block %dx%d unsupported\n", what, w, h); abort(); }
}
}  // namespace

#define UVGB_INTRA_LEAVES(SUF, PIX)                                                                              \
  extern "C" void uvg_b200_angular_pred_##SUF(int w, int h, int intra_mode, int channel_type, const PIX *above,   \
                                              const PIX *left, PIX *dst, int multi_ref_idx, int isp_mode)         \
  {                                                                                                              \
    leaf_check(w, h, "angular_pred");                                                                            \
    if (intra_mode < -14 || intra_mode > 80 || multi_ref_idx < 0 || multi_ref_idx > 3) {                          \
      fprintf(stderr, "uvg266_b200: angular_pred: mode %d / ref line %d out of range\n", intra_mode, multi_ref_idx); abort(); \
    }                                                                                                            \
    const LeafParams lp{ 0, w, h, intra_mode, channel_type, multi_ref_idx, isp_mode, (1 << dropin_bitdepth(sizeof(PIX))) - 1, kRefLen }; \
    leaf_call<PIX>(lp, above, left, dst, false);                                                                 \
  }                                                                                                              \
  extern "C" void uvg_b200_intra_pred_planar_##SUF(int w, int h, const PIX *top, const PIX *left, PIX *dst)       \
  {                                                                                                              \
    leaf_check(w, h, "intra_pred_planar");                                                                       \
    const LeafParams lp{ 1, w, h, 0, 0, 0, 0, (1 << dropin_bitdepth(sizeof(PIX))) - 1, kRefLen };                 \
    leaf_call<PIX>(lp, top, left, dst, false);                                                                   \
  }                                                                                                              \
  extern "C" void uvg_b200_pdpc_planar_dc_##SUF(int mode, int w, int h, const PIX *top, const PIX *left, PIX *dst) \
  {                                                                                                              \
    leaf_check(w, h, "pdpc_planar_dc");                                                                          \
    const LeafParams lp{ 2, w, h, mode, 0, 0, 0, (1 << dropin_bitdepth(sizeof(PIX))) - 1, kRefLen };              \
    leaf_call<PIX>(lp, top, left, dst, true);                                                                    \
  }
UVGB_INTRA_LEAVES(8, uint8_t)
UVGB_INTRA_LEAVES(16, uint16_t)
