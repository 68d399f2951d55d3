// picture.cu - batched block-distortion kernels (SAD / SATD / SSD / residual / MV search)
// for sm_100a. Semantics follow uvg266's generic strategies bit for bit
// (src/strategies/generic/picture-generic.c; per-function citations below).
//
// These kernels are HBM-bound byte/integer work: no tensor cores, the design points are
// coalesced 8/16-byte loads, enough independent loads in flight per thread, packed 16-bit
// butterflies for the 8-bit Hadamard, and grids sized in multiples of the SM count.
#include "common.cuh"
#include "hadamard.cuh"

namespace uvgb {

int launch_rough_umma(const Blocks &A, const Blocks &B, uint32_t *out, size_t n_items, int k, int n, cudaStream_t st);
int launch_rough_tma(const Blocks &A, const Blocks &B, uint32_t *out, size_t n_items, int k, int n, cudaStream_t st);
int launch_rough_mma(const Blocks &A, const Blocks &B, uint32_t *out, size_t n_items, int k, int n, cudaStream_t st);
int launch_rough_staged(const Blocks &A, const Blocks &B, uint32_t *out, size_t n_items, int k, int n,
                        cudaStream_t st);
int launch_sad_search_u8(const uint8_t *cur, const uint8_t *ref, int pic_w, int pic_h, int stride,
                         const int32_t *cur_x, const int32_t *cur_y, const int32_t *mv_x, const int32_t *mv_y,
                         int range, int w, int h, uint32_t *out, size_t n_blocks, cudaStream_t st);

// -------------------------------------------------------------------------------------
// pixel loads
// -------------------------------------------------------------------------------------
template <typename Pix> struct PixTraits;
template <> struct PixTraits<uint8_t>  { static constexpr int kPerWord = 4; };
template <> struct PixTraits<uint16_t> { static constexpr int kPerWord = 2; };

// Load N pixels (N*sizeof(Pix) in {4, 8, 16} bytes) into 32-bit words, using one vector
// load when the address allows it and narrower loads otherwise (MV-displaced blocks are
// arbitrarily aligned).
template <typename Pix, int NWORDS>
__device__ __forceinline__ void load_words(const Pix *p, uint32_t (&w)[NWORDS])
{
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  if constexpr (NWORDS == 4) {
    if ((a & 15) == 0) {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
      w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
      return;
    }
  }
  if constexpr (NWORDS >= 2) {
    if ((a & 7) == 0) {
#pragma unroll
      for (int i = 0; i < NWORDS; i += 2) {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p) + i / 2);
        w[i] = v.x; w[i + 1] = v.y;
      }
      return;
    }
  }
  if ((a & 3) == 0) {
#pragma unroll
    for (int i = 0; i < NWORDS; ++i) w[i] = __ldg(reinterpret_cast<const uint32_t *>(p) + i);
  } else {
    constexpr int PW = PixTraits<Pix>::kPerWord;
#pragma unroll
    for (int i = 0; i < NWORDS; ++i) {
      uint32_t v = 0;
#pragma unroll
      for (int j = 0; j < PW; ++j) v |= (uint32_t)__ldg(p + i * PW + j) << (j * (32 / PW));
      w[i] = v;
    }
  }
}

// -------------------------------------------------------------------------------------
// SAD / SSD
// -------------------------------------------------------------------------------------
enum { OP_SAD = 0, OP_SSD = 1 };

template <typename Pix, int OP>
__device__ __forceinline__ uint32_t acc_word(uint32_t a, uint32_t b, uint32_t acc)
{
  if (sizeof(Pix) == 1) {
    if (OP == OP_SAD) return __vsadu4(a, b) + acc;
    const uint32_t d = __vabsdiffu4(a, b);
    return __dp4a(d, d, acc);
  } else {
    const uint32_t d = __vabsdiffu2(a, b);
    const uint32_t lo = d & 0xffffu, hi = d >> 16;
    if (OP == OP_SAD) return acc + lo + hi;
    return acc + lo * lo + hi * hi;
  }
}

// One item = (block i of a, block i*k+c of b). G lanes cooperate on an item; each lane
// walks 4-word (16-byte) or 1-word chunks of the rows. w*sizeof(Pix) is a multiple of 4.
template <typename Pix, int OP, int G, int NWORDS>
__global__ void __launch_bounds__(256)
sad_ssd_kernel(Blocks A, Blocks B, uint32_t *__restrict__ out, size_t n_items, int k, int w, int h,
               int shift)
{
  constexpr int PW = PixTraits<Pix>::kPerWord;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t item = gid / G;
  const int lane = (int)(gid % G);
  if (item >= n_items) return;  // whole groups leave together (G divides 32)
  const Pix *pa = static_cast<const Pix *>(A.base) + A.offset(item / k);
  const Pix *pb = static_cast<const Pix *>(B.base) + B.offset(item);
  const int chunks_per_row = w / (PW * NWORDS);
  const int n_chunks = chunks_per_row * h;
  uint32_t acc = 0;
  for (int c = lane; c < n_chunks; c += G) {
    const int y = c / chunks_per_row, x = (c - y * chunks_per_row) * (PW * NWORDS);
    uint32_t wa[NWORDS], wb[NWORDS];
    load_words<Pix, NWORDS>(pa + (size_t)y * A.stride + x, wa);
    load_words<Pix, NWORDS>(pb + (size_t)y * B.stride + x, wb);
#pragma unroll
    for (int i = 0; i < NWORDS; ++i) acc = acc_word<Pix, OP>(wa[i], wb[i], acc);
  }
  acc = group_sum<G>(acc);
  if (lane == 0) {
    // pixels_calc_ssd accumulates in int and shifts arithmetically (picture-generic.c:1129)
    out[item] = (OP == OP_SSD) ? (uint32_t)((int32_t)acc >> shift) : (acc >> shift);
  }
}

// Arbitrary widths (w % 4 != 0): the frame-edge splitter image_interpolated_sad (image.c:310-427)
// calls reg_sad / hor_sad / ver_sad on slivers such as 32x1 or 3x8. Pixel-at-a-time path.
template <typename Pix, int OP>
__global__ void __launch_bounds__(128)
sad_ssd_px_kernel(Blocks A, Blocks B, uint32_t *__restrict__ out, size_t n_items, int k, int w, int h,
                  int shift)
{
  const size_t item = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one warp per item
  const int lane = threadIdx.x & 31;
  if (item >= n_items) return;
  const Pix *pa = static_cast<const Pix *>(A.base) + A.offset(item / k);
  const Pix *pb = static_cast<const Pix *>(B.base) + B.offset(item);
  uint32_t acc = 0;
  for (int e = lane; e < w * h; e += 32) {
    const int y = e / w, x = e - y * w;
    const int d = (int)__ldg(pa + (size_t)y * A.stride + x) - (int)__ldg(pb + (size_t)y * B.stride + x);
    acc += (OP == OP_SAD) ? (uint32_t)abs(d) : (uint32_t)(d * d);
  }
  acc = group_sum<32>(acc);
  if (lane == 0) out[item] = (OP == OP_SSD) ? (uint32_t)((int32_t)acc >> shift) : (acc >> shift);
}

// -------------------------------------------------------------------------------------
// Hadamard building blocks (32-bit path: any bit depth)
// -------------------------------------------------------------------------------------
// In-register W x H Walsh-Hadamard of differences; returns sum|coef| with the DC term
// replaced by |dc|>>2 (hadamard_4x4_generic :195-197, satd_8x8_subblock_generic :342-343,
// xCalcHADs* :645,784,860,941). Output order is irrelevant for the sum.
template <int W, int H>
__device__ __forceinline__ int had_abs_sum(int (&d)[W * H])
{
#pragma unroll
  for (int y = 0; y < H; ++y)
#pragma unroll
    for (int len = 1; len < W; len <<= 1)
#pragma unroll
      for (int i = 0; i < W; i += len << 1)
#pragma unroll
        for (int j = i; j < i + len; ++j) {
          const int a = d[y * W + j], b = d[y * W + j + len];
          d[y * W + j] = a + b; d[y * W + j + len] = a - b;
        }
#pragma unroll
  for (int x = 0; x < W; ++x)
#pragma unroll
    for (int len = 1; len < H; len <<= 1)
#pragma unroll
      for (int i = 0; i < H; i += len << 1)
#pragma unroll
        for (int j = i; j < i + len; ++j) {
          const int a = d[j * W + x], b = d[(j + len) * W + x];
          d[j * W + x] = a + b; d[(j + len) * W + x] = a - b;
        }
  int s = 0;
#pragma unroll
  for (int i = 1; i < W * H; ++i) s += abs(d[i]);
  s += abs(d[0]) >> 2;
  return s;
}

template <typename Pix, int W, int H>
__device__ __forceinline__ void load_diff(const Pix *o, int so, const Pix *c, int sc, int (&d)[W * H])
{
  constexpr int PW = PixTraits<Pix>::kPerWord;
  constexpr int NW = (W / PW) > 4 ? 4 : (W / PW);       // words per load (<= 16 bytes)
  constexpr int LOADS = (W / PW) / NW;
#pragma unroll
  for (int y = 0; y < H; ++y) {
#pragma unroll
    for (int l = 0; l < LOADS; ++l) {
      uint32_t wo[NW], wc[NW];
      load_words<Pix, NW>(o + (size_t)y * so + l * NW * PW, wo);
      load_words<Pix, NW>(c + (size_t)y * sc + l * NW * PW, wc);
#pragma unroll
      for (int i = 0; i < NW; ++i)
#pragma unroll
        for (int j = 0; j < PW; ++j) {
          const int sh = j * (32 / PW);
          const int mask = (1 << (32 / PW)) - 1;
          d[y * W + (l * NW + i) * PW + j] = (int)((wo[i] >> sh) & mask) - (int)((wc[i] >> sh) & mask);
        }
    }
  }
}

template <typename Pix>
__device__ __forceinline__ unsigned satd4_sub(const Pix *o, int so, const Pix *c, int sc)
{
  int d[16];
  load_diff<Pix, 4, 4>(o, so, c, sc, d);
  return (unsigned)((had_abs_sum<4, 4>(d) + 1) >> 1);          // :198
}

// ---- 8x8 sub-block ------------------------------------------------------------------
// 8-bit fast path: packed 16-bit butterflies, see hadamard.cuh.
__device__ __forceinline__ unsigned satd8_sub_u8(const uint8_t *o, int so, const uint8_t *c, int sc)
{
  uint32_t wo[8][2], wc[8][2];
#pragma unroll
  for (int y = 0; y < 8; ++y) {
    load_words<uint8_t, 2>(o + (size_t)y * so, wo[y]);
    load_words<uint8_t, 2>(c + (size_t)y * sc, wc[y]);
  }
  unsigned sad;
  return satd8_words_u8<false>(wo, wc, sad);
}

template <typename Pix>
__device__ __forceinline__ unsigned satd8_sub(const Pix *o, int so, const Pix *c, int sc)
{
  if (sizeof(Pix) == 1) {
    return satd8_sub_u8(reinterpret_cast<const uint8_t *>(o), so, reinterpret_cast<const uint8_t *>(c), sc);
  } else {
    int d[64];
    load_diff<Pix, 8, 8>(o, so, c, sc, d);
    return (unsigned)((had_abs_sum<8, 8>(d) + 2) >> 2);
  }
}

// Sub-block decomposition of a w*h block, computed on the host per call.
//   n4c 4x4 blocks down column 0; n4r 4x4 blocks along row 0 starting at x4r; then an
//   n8x * n8y grid of 8x8 blocks starting at (x8, y8).
// SATD_ANY_SIZE (strategies-picture.h:75-113) and SATD_ANY_SIZE_MULTI_GENERIC
// (picture-generic.c:412-479, literal edge handling) differ only in x4r / y8.
struct SatdLayout { int n4c, n4r, x4r, n8x, n8y, x8, y8; };

static SatdLayout make_satd_layout(int w, int h, bool quad)
{
  SatdLayout L{};
  const int wm8 = w % 8;
  int ww = w, hh = h;
  if (wm8) { L.n4c = h / 4; ww -= 4; }
  if (h % 8) {
    L.n4r = ww / 4;
    L.x4r = quad ? 0 : (wm8 ? 4 : 0);
    hh -= 4;
  }
  if (quad) {
    L.x8 = wm8; L.y8 = hh % 8;
    L.n8x = (ww - L.x8 + 7) / 8; if (L.n8x < 0) L.n8x = 0;
    L.n8y = (hh - L.y8 + 7) / 8; if (L.n8y < 0) L.n8y = 0;
  } else {
    L.x8 = wm8 ? 4 : 0; L.y8 = (h % 8) ? 4 : 0;
    L.n8x = ww / 8; L.n8y = hh / 8;
  }
  return L;
}

template <typename Pix, int G>
__global__ void __launch_bounds__(128)
satd_kernel(Blocks A, Blocks B, uint32_t *__restrict__ out, size_t n_items, int k, SatdLayout L,
            int shift)
{
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t item = gid / G;
  const int lane = (int)(gid % G);
  if (item >= n_items) return;
  const Pix *pa = static_cast<const Pix *>(A.base) + A.offset(item / k);
  const Pix *pb = static_cast<const Pix *>(B.base) + B.offset(item);
  const int sa = (int)A.stride, sb = (int)B.stride;
  const int n8 = L.n8x * L.n8y;
  const int nsub = L.n4c + L.n4r + n8;
  unsigned acc = 0;
  for (int s = lane; s < nsub; s += G) {
    if (s < n8) {
      const int by = s / L.n8x, bx = s - by * L.n8x;
      const int x = L.x8 + bx * 8, y = L.y8 + by * 8;
      acc += satd8_sub<Pix>(pa + (size_t)y * sa + x, sa, pb + (size_t)y * sb + x, sb);
    } else if (s < n8 + L.n4c) {
      const int y = (s - n8) * 4;
      acc += satd4_sub<Pix>(pa + (size_t)y * sa, sa, pb + (size_t)y * sb, sb);
    } else {
      const int x = L.x4r + (s - n8 - L.n4c) * 4;
      acc += satd4_sub<Pix>(pa + x, sa, pb + x, sb);
    }
  }
  acc = group_sum<G>(acc);
  if (lane == 0) out[item] = acc >> shift;
}

// ---- VTM-style Hadamard (xGetHADs, picture-generic.c:949-1049) -----------------------
template <typename Pix, int W, int H>
__device__ __forceinline__ unsigned had_rect(const Pix *o, int so, const Pix *c, int sc)
{
  int d[W * H];
  load_diff<Pix, W, H>(o, so, c, sc, d);
  const int s = had_abs_sum<W, H>(d);
  // sad = (int)(sad / sqrt(W*H) * 2) in IEEE double, round-to-nearest divide (:647,786,862,943)
  const double norm = (W * H == 128) ? 11.313708498984761 /* sqrt(16.0*8) */
                                     : 5.656854249492381  /* sqrt(4.0*8)  */;
  return (unsigned)(int)(__ddiv_rn((double)s, norm) * 2.0);
}

template <typename Pix>
__device__ __forceinline__ unsigned had2x2(const Pix *o, int so, const Pix *c, int sc)
{
  // xCalcHADs2x2 :481-501 - int16 intermediates
  const short d0 = (short)((int)o[0] - (int)c[0]), d1 = (short)((int)o[1] - (int)c[1]);
  const short d2 = (short)((int)o[so] - (int)c[sc]), d3 = (short)((int)o[so + 1] - (int)c[sc + 1]);
  const short m0 = (short)(d0 + d2), m1 = (short)(d1 + d3), m2 = (short)(d0 - d2), m3 = (short)(d1 - d3);
  return (unsigned)((abs(m0 + m1) >> 2) + abs(m0 - m1) + abs(m2 + m3) + abs(m2 - m3));
}

// kind: 0 16x8, 1 8x16, 2 8x4, 3 4x8, 4 8x8, 5 4x4, 6 2x2
template <typename Pix, int G>
__global__ void __launch_bounds__(64)
satd_vtm_kernel(Blocks A, Blocks B, uint32_t *__restrict__ out, size_t n_items, int k, int w, int h,
                int kind)
{
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t item = gid / G;
  const int lane = (int)(gid % G);
  if (item >= n_items) return;
  const Pix *pa = static_cast<const Pix *>(A.base) + A.offset(item / k);
  const Pix *pb = static_cast<const Pix *>(B.base) + B.offset(item);
  const int sa = (int)A.stride, sb = (int)B.stride;
  const int bw = kind == 0 ? 16 : (kind == 1 || kind == 2 || kind == 4) ? 8 : (kind == 6 ? 2 : 4);
  const int bh = kind == 1 ? 16 : (kind == 0 || kind == 3 || kind == 4) ? 8 : (kind == 6 ? 2 : 4);
  const int nx = w / bw, nsub = nx * (h / bh);
  unsigned acc = 0;  // the reference sums in uint64 and truncates to unsigned: same low 32 bits
  for (int s = lane; s < nsub; s += G) {
    const int by = s / nx, bx = s - by * nx;
    const Pix *o = pa + (size_t)(by * bh) * sa + bx * bw;
    const Pix *c = pb + (size_t)(by * bh) * sb + bx * bw;
    switch (kind) {
      case 0: acc += had_rect<Pix, 16, 8>(o, sa, c, sb); break;
      case 1: acc += had_rect<Pix, 8, 16>(o, sa, c, sb); break;
      case 2: acc += had_rect<Pix, 8, 4>(o, sa, c, sb); break;
      case 3: acc += had_rect<Pix, 4, 8>(o, sa, c, sb); break;
      case 4: acc += satd8_sub<Pix>(o, sa, c, sb); break;
      case 5: acc += satd4_sub<Pix>(o, sa, c, sb); break;
      default: acc += had2x2<Pix>(o, sa, c, sb); break;
    }
  }
  acc = group_sum<G>(acc);
  if (lane == 0) out[item] = acc;
}

// -------------------------------------------------------------------------------------
// Intra rough search cost: min(SATD, 2*SAD) of square n*n blocks, one original against k
// candidate predictions (get_cost_dual, search_intra.c:133-192 on top of
// satd_NxN_dual :371-410 and sad_NxN_dual :1091-1113).
// -------------------------------------------------------------------------------------
template <typename Pix, int G>
__global__ void __launch_bounds__(128)
rough_cost_kernel(Blocks A, Blocks B, uint32_t *__restrict__ out, size_t n_items, int k, int n,
                  int shift)
{
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t item = gid / G;
  const int lane = (int)(gid % G);
  if (item >= n_items) return;
  const Pix *pa = static_cast<const Pix *>(A.base) + A.offset(item / k);
  const Pix *pb = static_cast<const Pix *>(B.base) + B.offset(item);
  const int sa = (int)A.stride, sb = (int)B.stride;
  unsigned satd = 0, sad = 0;
  if (n == 4) {
    if (lane == 0) {
      int d[16];
      load_diff<Pix, 4, 4>(pa, sa, pb, sb, d);
#pragma unroll
      for (int i = 0; i < 16; ++i) sad += abs(d[i]);
      satd = (unsigned)((had_abs_sum<4, 4>(d) + 1) >> 1);
    }
    satd = group_sum<G>(satd); sad = group_sum<G>(sad);
    // satd_4x4_dual has no bit-depth shift (:401-405); sad_4x4_dual has (:1091)
    if (lane == 0) out[item] = min(satd, 2 * (sad >> shift));
    return;
  }
  const int nb = n / 8, nsub = nb * nb;
  constexpr int PW = PixTraits<Pix>::kPerWord;
  for (int s = lane; s < nsub; s += G) {
    const int by = s / nb, bx = s - by * nb;
    const Pix *o = pa + (size_t)(by * 8) * sa + bx * 8;
    const Pix *c = pb + (size_t)(by * 8) * sb + bx * 8;
    satd += satd8_sub<Pix>(o, sa, c, sb);
    // the same 8x8 pixels again for the SAD (L1 hits; the loads are shared by the compiler
    // where it can prove it)
#pragma unroll
    for (int y = 0; y < 8; ++y) {
      uint32_t wo[8 / PW], wc[8 / PW];
      load_words<Pix, 8 / PW>(o + (size_t)y * sa, wo);
      load_words<Pix, 8 / PW>(c + (size_t)y * sb, wc);
#pragma unroll
      for (int i = 0; i < 8 / PW; ++i) sad = acc_word<Pix, OP_SAD>(wo[i], wc[i], sad);
    }
  }
  satd = group_sum<G>(satd); sad = group_sum<G>(sad);
  if (lane == 0) out[item] = min(satd >> shift, 2 * (sad >> shift));
}

// -------------------------------------------------------------------------------------
// residual = ref - pred (generate_residual_generic :1360-1369), dense int16 output
// -------------------------------------------------------------------------------------
template <typename Pix>
__global__ void __launch_bounds__(256)
residual_kernel(Blocks A, Blocks B, int16_t *__restrict__ out, size_t n_blocks, int w, int h)
{
  const size_t total4 = n_blocks * (size_t)(w * h / 4);
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total4;
       t += (size_t)gridDim.x * blockDim.x) {
    const size_t blk = t / (w * h / 4);
    const int e = (int)(t - blk * (w * h / 4)) * 4;
    const int y = e / w, x = e - y * w;
    const Pix *pa = static_cast<const Pix *>(A.base) + A.offset(blk) + (size_t)y * A.stride + x;
    const Pix *pb = static_cast<const Pix *>(B.base) + B.offset(blk) + (size_t)y * B.stride + x;
    short r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = (short)((int)__ldg(pa + i) - (int)__ldg(pb + i));
    *reinterpret_cast<uint2 *>(out + blk * (size_t)(w * h) + e) =
        make_uint2((uint16_t)r[0] | ((uint32_t)(uint16_t)r[1] << 16),
                   (uint16_t)r[2] | ((uint32_t)(uint16_t)r[3] << 16));
  }
}

// -------------------------------------------------------------------------------------
// Motion-search SAD over a (2r+1)^2 window with frame-edge replication.
// One CTA per current block: the (w+2r) x (h+2r) reference window is staged once in
// shared memory (clamped coordinates == ver_sad/hor_sad/cor_sad of image.c:280-427), the
// current block too; every thread then evaluates candidates from shared memory.
// -------------------------------------------------------------------------------------
template <typename Pix>
__global__ void __launch_bounds__(256)
sad_search_kernel(const Pix *__restrict__ cur, const Pix *__restrict__ ref, int pic_w, int pic_h,
                  int stride, const int32_t *__restrict__ cur_x, const int32_t *__restrict__ cur_y,
                  const int32_t *__restrict__ mv_x, const int32_t *__restrict__ mv_y, int range,
                  int w, int h, uint32_t *__restrict__ out, int shift)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ww = w + 2 * range, wh = h + 2 * range;
  const int wpitch = ww + 1;  // +1 pixel: neighbouring candidates start on different banks
  Pix *s_ref = reinterpret_cast<Pix *>(smem_raw);
  Pix *s_cur = s_ref + (size_t)wpitch * wh;
  const size_t j = blockIdx.x;
  const int cx = cur_x[j], cy = cur_y[j];
  const int ox = cx + mv_x[j] - range, oy = cy + mv_y[j] - range;
  for (int t = threadIdx.x; t < ww * wh; t += blockDim.x) {
    const int y = t / ww, x = t - y * ww;
    const int rx = clip3(0, pic_w - 1, ox + x), ry = clip3(0, pic_h - 1, oy + y);
    s_ref[y * wpitch + x] = __ldg(ref + (size_t)ry * stride + rx);
  }
  for (int t = threadIdx.x; t < w * h; t += blockDim.x) {
    const int y = t / w, x = t - y * w;
    s_cur[t] = __ldg(cur + (size_t)(cy + y) * stride + cx + x);
  }
  __syncthreads();
  const int side = 2 * range + 1, ncand = side * side;
  for (int c = threadIdx.x; c < ncand; c += blockDim.x) {
    const int dy = c / side, dx = c - dy * side;
    unsigned acc = 0;
    for (int y = 0; y < h; ++y) {
      const Pix *r = s_ref + (dy + y) * wpitch + dx;
      const Pix *p = s_cur + y * w;
#pragma unroll 4
      for (int x = 0; x < w; ++x) acc += (unsigned)abs((int)p[x] - (int)r[x]);
    }
    out[j * (size_t)ncand + c] = acc >> shift;
  }
}

// -------------------------------------------------------------------------------------
// launchers
// -------------------------------------------------------------------------------------
static inline unsigned grid_for(size_t threads, int block)
{
  size_t g = (threads + block - 1) / block;
  return (unsigned)(g ? g : 1);
}

static int pick_group(int work_units)
{
  int g = 1;
  while (g < work_units && g < 32) g <<= 1;
  return g;
}

template <typename Pix, int OP>
static int launch_sad_ssd(const Blocks &A, const Blocks &B, uint32_t *out, size_t n_items, int k,
                          int w, int h, int shift, cudaStream_t st)
{
  constexpr int PW = PixTraits<Pix>::kPerWord;
  if (w % PW) {
    sad_ssd_px_kernel<Pix, OP><<<grid_for(n_items * 32, 128), 128, 0, st>>>(A, B, out, n_items, k, w, h, shift);
    UVGB_LAUNCH_CHECK();
    return 0;
  }
  // Small dense blocks (4x4, 8-bit 8x8: the whole block is <= 64 contiguous bytes): treat the block as one
  // row of w*h pixels and give every lane ONE 16-byte chunk, so that a warp's loads are fully coalesced
  // (as rows, an 8x8 u8 block is sixteen 4-byte chunks per operand: 41 % of the HBM roofline).
  const size_t blk_bytes = (size_t)w * h * sizeof(Pix);
  const bool linear = A.stride == (uint32_t)w && B.stride == (uint32_t)w && blk_bytes <= 64 && blk_bytes % 16 == 0 && h > 1;
  if (linear) { w *= h; h = 1; }
  const bool wide = (w % (4 * PW)) == 0;
  const int chunks = wide ? (w / (4 * PW)) * h : (w / PW) * h;
  // aim for >= 4 chunks per lane so several loads are in flight per thread
  int G = linear ? pick_group(chunks) : pick_group((chunks + 3) / 4);
  const int block = 256;
#define UVGB_SAD_CASE(GG)                                                                      \
  case GG:                                                                                     \
    if (wide) sad_ssd_kernel<Pix, OP, GG, 4><<<grid_for(n_items * GG, block), block, 0, st>>>( \
        A, B, out, n_items, k, w, h, shift);                                                   \
    else sad_ssd_kernel<Pix, OP, GG, 1><<<grid_for(n_items * GG, block), block, 0, st>>>(      \
        A, B, out, n_items, k, w, h, shift);                                                   \
    break;
  switch (G) {
    UVGB_SAD_CASE(1) UVGB_SAD_CASE(2) UVGB_SAD_CASE(4) UVGB_SAD_CASE(8) UVGB_SAD_CASE(16)
    default: UVGB_SAD_CASE(32)
  }
#undef UVGB_SAD_CASE
  UVGB_LAUNCH_CHECK();
  return 0;
}

template <typename Pix>
static int launch_satd(const Blocks &A, const Blocks &B, uint32_t *out, size_t n_items, int k,
                       const SatdLayout &L, int shift, cudaStream_t st)
{
  const int nsub = L.n4c + L.n4r + L.n8x * L.n8y;
  const int G = pick_group(nsub);
  const int block = 128;
#define UVGB_SATD_CASE(GG)                                                                        \
  case GG:                                                                                        \
    satd_kernel<Pix, GG><<<grid_for(n_items * GG, block), block, 0, st>>>(A, B, out, n_items, k, \
                                                                            L, shift);            \
    break;
  switch (G) {
    UVGB_SATD_CASE(1) UVGB_SATD_CASE(2) UVGB_SATD_CASE(4) UVGB_SATD_CASE(8) UVGB_SATD_CASE(16)
    default: UVGB_SATD_CASE(32)
  }
#undef UVGB_SATD_CASE
  UVGB_LAUNCH_CHECK();
  return 0;
}

template <typename Pix>
static int launch_vtm(const Blocks &A, const Blocks &B, uint32_t *out, size_t n_items, int k, int w,
                      int h, cudaStream_t st)
{
  int kind, bw, bh;
  if (w > h && (h & 7) == 0 && (w & 15) == 0) { kind = 0; bw = 16; bh = 8; }
  else if (w < h && (w & 7) == 0 && (h & 15) == 0) { kind = 1; bw = 8; bh = 16; }
  else if (w > h && (h & 3) == 0 && (w & 7) == 0) { kind = 2; bw = 8; bh = 4; }
  else if (w < h && (w & 3) == 0 && (h & 7) == 0) { kind = 3; bw = 4; bh = 8; }
  else if (h % 8 == 0 && w % 8 == 0) { kind = 4; bw = 8; bh = 8; }
  else if (h % 4 == 0 && w % 4 == 0) { kind = 5; bw = 4; bh = 4; }
  else if (h % 2 == 0 && w % 2 == 0) { kind = 6; bw = 2; bh = 2; }
  else { return cudaMemsetAsync(out, 0, n_items * sizeof(uint32_t), st) != cudaSuccess; }
  const int nsub = (w / bw) * (h / bh);
  const int G = pick_group(nsub);
  const int block = 64;
#define UVGB_VTM_CASE(GG)                                                                      \
  case GG:                                                                                     \
    satd_vtm_kernel<Pix, GG><<<grid_for(n_items * GG, block), block, 0, st>>>(A, B, out,       \
                                                                                n_items, k, w, \
                                                                                h, kind);      \
    break;
  switch (G) {
    UVGB_VTM_CASE(1) UVGB_VTM_CASE(2) UVGB_VTM_CASE(4) UVGB_VTM_CASE(8) UVGB_VTM_CASE(16)
    default: UVGB_VTM_CASE(32)
  }
#undef UVGB_VTM_CASE
  UVGB_LAUNCH_CHECK();
  return 0;
}

template <typename Pix>
static int distortion_impl(int metric, int bitdepth, int w, int h, const Blocks &A, const Blocks &B,
                           uint32_t *out, size_t n_blocks, int k, cudaStream_t st)
{
  const size_t n_items = n_blocks * (size_t)k;
  const int bdshift = bitdepth - 8;
  switch (metric) {
    case UVG_B200_SAD: return launch_sad_ssd<Pix, OP_SAD>(A, B, out, n_items, k, w, h, 0, st);
    case UVG_B200_SAD_NXN: return launch_sad_ssd<Pix, OP_SAD>(A, B, out, n_items, k, w, h, bdshift, st);
    case UVG_B200_SSD: return launch_sad_ssd<Pix, OP_SSD>(A, B, out, n_items, k, w, h, 2 * bdshift, st);
    case UVG_B200_SATD: return launch_satd<Pix>(A, B, out, n_items, k, make_satd_layout(w, h, false), bdshift, st);
    case UVG_B200_SATD_QUAD: return launch_satd<Pix>(A, B, out, n_items, k, make_satd_layout(w, h, true), bdshift, st);
    case UVG_B200_SATD_4X4: {
      SatdLayout L{}; L.n4c = 1;  // one 4x4 at (0,0), no shift (satd_4x4_generic :205)
      return launch_satd<Pix>(A, B, out, n_items, k, L, 0, st);
    }
    case UVG_B200_SATD_VTM: return launch_vtm<Pix>(A, B, out, n_items, k, w, h, st);
    default: set_error("unknown metric %d", metric); return 1;
  }
}

static int check_dims(int bitdepth, int w, int h, int k)
{
  if (bitdepth < 8 || bitdepth > 16) { set_error("bitdepth %d out of range", bitdepth); return 1; }
  if (w <= 0 || h <= 0 || w > 128 || h > 128) { set_error("block %dx%d out of range", w, h); return 1; }
  if (k < 1) { set_error("k must be >= 1"); return 1; }
  return 0;
}

}  // namespace uvgb

using namespace uvgb;

extern "C" int uvg_b200_distortion_dev(int metric, int bitdepth, int w, int h,
                                       const uvg_b200_blocks *a, const uvg_b200_blocks *b,
                                       uint32_t *out, size_t n_blocks, int k, void *stream)
{
  if (ensure_init()) return 1;
  if (check_dims(bitdepth, w, h, k)) return 1;
  if (n_blocks == 0) return 0;
  const bool is_satd = metric == UVG_B200_SATD || metric == UVG_B200_SATD_4X4 || metric == UVG_B200_SATD_QUAD;
  if (is_satd && ((w % 4) || (h % 4))) {
    set_error("metric %d needs w,h multiples of 4 (got %dx%d)", metric, w, h);
    return 1;
  }
  if (metric == UVG_B200_SATD_VTM && bitdepth > 8 && (w % 2)) { set_error("odd width"); return 1; }
  const Blocks A = to_blocks(a), B = to_blocks(b);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (bitdepth == 8) return distortion_impl<uint8_t>(metric, bitdepth, w, h, A, B, out, n_blocks, k, st);
  return distortion_impl<uint16_t>(metric, bitdepth, w, h, A, B, out, n_blocks, k, st);
}

extern "C" int uvg_b200_rough_cost_dev(int bitdepth, int n, const uvg_b200_blocks *orig,
                                       const uvg_b200_blocks *preds, uint32_t *out,
                                       size_t n_blocks, int k, void *stream)
{
  if (ensure_init()) return 1;
  if (check_dims(bitdepth, n, n, k)) return 1;
  if (n != 4 && n != 8 && n != 16 && n != 32 && n != 64) { set_error("rough cost: n=%d", n); return 1; }
  if (n_blocks == 0) return 0;
  const Blocks A = to_blocks(orig), B = to_blocks(preds);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n_items = n_blocks * (size_t)k;
  if (bitdepth == 8) {
    // dense candidates, interchangeable kernels (all bit-exact, all in the GPU tests); one 4K frame of the bench
    // workload on a B200, CUDA events:                                                     8x8 k=23   16x16 k=24
    //   tc5 (default where it applies: dense originals too, n = 8 / 16, k >= 10)
    //                     tcgen05.mma kind::i8 fed by tensor-map TMA, H O H^T from the same
    //                     MMA, accumulators in TMEM (rough_tma.cu)                          0.068 ms   0.060 ms
    //   ROUGH=staged      packed-ALU shared-memory staged kernel (rough.cu)                 0.072 ms   0.081 ms
    //   ROUGH=mma         Hadamard as mma.sync IMMA (rough_mma.cu)                          0.090 ms   0.070 ms
    //   ROUGH=umma        round 1's tcgen05 kernel fed by per-thread cp.async (rough_umma.cu) 0.082 ms 0.107 ms
    //   ROUGH=tma         the staged kernel fed by cp.async.bulk                            0.081 ms   0.083 ms
    // Without ROUGH the fastest measured kernel that accepts the layout runs: tc5, else staged (8x8) / mma (>= 16x16).
    static const char *mode = getenv("UVG266_B200_ROUGH");
    if (!mode || !strcmp(mode, "tc5")) {
      const int rt = launch_rough_tma(A, B, out, n_items, k, n, st);
      if (rt >= 0) return rt;
    }
    const bool want_umma = mode && !strcmp(mode, "umma");
    const bool want_mma = mode ? !strcmp(mode, "mma") : n >= 16;
    if (want_umma) {
      const int ru = launch_rough_umma(A, B, out, n_items, k, n, st);
      if (ru >= 0) return ru;
    }
    if (want_mma) {
      const int rm = launch_rough_mma(A, B, out, n_items, k, n, st);
      if (rm >= 0) return rm;
    }
    const int rc = launch_rough_staged(A, B, out, n_items, k, n, st);
    if (rc >= 0) return rc;
  }
  const int nsub = n == 4 ? 1 : (n / 8) * (n / 8);
  const int G = pick_group(nsub);
  const int block = 128;
  const int shift = bitdepth - 8;
#define UVGB_RC_CASE(PIX, GG)                                                                    \
  case GG:                                                                                       \
    rough_cost_kernel<PIX, GG><<<grid_for(n_items * GG, block), block, 0, st>>>(A, B, out,      \
                                                                                 n_items, k, n, \
                                                                                 shift);        \
    break;
  if (bitdepth == 8) {
    switch (G) { UVGB_RC_CASE(uint8_t, 1) UVGB_RC_CASE(uint8_t, 2) UVGB_RC_CASE(uint8_t, 4)
                 UVGB_RC_CASE(uint8_t, 8) UVGB_RC_CASE(uint8_t, 16) default: UVGB_RC_CASE(uint8_t, 32) }
  } else {
    switch (G) { UVGB_RC_CASE(uint16_t, 1) UVGB_RC_CASE(uint16_t, 2) UVGB_RC_CASE(uint16_t, 4)
                 UVGB_RC_CASE(uint16_t, 8) UVGB_RC_CASE(uint16_t, 16) default: UVGB_RC_CASE(uint16_t, 32) }
  }
#undef UVGB_RC_CASE
  UVGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int uvg_b200_residual_dev(int bitdepth, int w, int h, const uvg_b200_blocks *ref,
                                     const uvg_b200_blocks *pred, int16_t *residual,
                                     size_t n_blocks, void *stream)
{
  if (ensure_init()) return 1;
  if (check_dims(bitdepth, w, h, 1)) return 1;
  if ((w * h) % 4 || (w % 4)) { set_error("residual: width must be a multiple of 4"); return 1; }
  if (n_blocks == 0) return 0;
  const Blocks A = to_blocks(ref), B = to_blocks(pred);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t total4 = n_blocks * (size_t)(w * h / 4);
  const int block = 256;
  unsigned grid = grid_for(total4, block);
  const unsigned cap = (unsigned)sm_count() * 16;
  if (grid > cap) grid = cap;
  if (bitdepth == 8) residual_kernel<uint8_t><<<grid, block, 0, st>>>(A, B, residual, n_blocks, w, h);
  else residual_kernel<uint16_t><<<grid, block, 0, st>>>(A, B, residual, n_blocks, w, h);
  UVGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int uvg_b200_sad_search_dev(int bitdepth, int w, int h, const void *cur, const void *ref,
                                       int pic_w, int pic_h, int stride, const int32_t *cur_x,
                                       const int32_t *cur_y, const int32_t *mv_x,
                                       const int32_t *mv_y, int range, uint32_t *out,
                                       size_t n_blocks, void *stream)
{
  if (ensure_init()) return 1;
  if (check_dims(bitdepth, w, h, 1)) return 1;
  if (range < 0 || range > 64) { set_error("search range %d out of [0,64]", range); return 1; }
  if (n_blocks == 0) return 0;
  if (bitdepth == 8) {
    const int rc = launch_sad_search_u8(static_cast<const uint8_t *>(cur), static_cast<const uint8_t *>(ref), pic_w,
                                        pic_h, stride, cur_x, cur_y, mv_x, mv_y, range, w, h, out, n_blocks,
                                        static_cast<cudaStream_t>(stream));
    if (rc >= 0) return rc;
  }
  const int pix = bitdepth == 8 ? 1 : 2;
  const size_t smem = ((size_t)(w + 2 * range + 1) * (h + 2 * range) + (size_t)w * h) * pix;
  if (smem > 227 * 1024) { set_error("search window does not fit shared memory"); return 1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int shift = bitdepth - 8;
  if (bitdepth == 8) {
    UVGB_TRY(cudaFuncSetAttribute(sad_search_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sad_search_kernel<uint8_t><<<(unsigned)n_blocks, 256, smem, st>>>(
        static_cast<const uint8_t *>(cur), static_cast<const uint8_t *>(ref), pic_w, pic_h, stride,
        cur_x, cur_y, mv_x, mv_y, range, w, h, out, shift);
  } else {
    UVGB_TRY(cudaFuncSetAttribute(sad_search_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sad_search_kernel<uint16_t><<<(unsigned)n_blocks, 256, smem, st>>>(
        static_cast<const uint16_t *>(cur), static_cast<const uint16_t *>(ref), pic_w, pic_h, stride,
        cur_x, cur_y, mv_x, mv_y, range, w, h, out, shift);
  }
  UVGB_LAUNCH_CHECK();
  return 0;
}
