// frame.cu - device-resident all-intra frame pipeline behind the C ABI (uvg_b200_frame_*).
//
// The batched call site of SURVEY.md 8(f)-4: what the encoder does per CTU for an all-intra picture -
//   search_intra_rough          search_intra.c:986-1062   predict k candidate modes per block, cost = min(SATD, 2 SAD)
//   (mode selection)            search_intra.c:1064-1105  keep the best candidates
//   uvg_quantize_lcu_residual   transform.c:1487-1600     predict with the chosen mode -> quantize_residual ->
//                                                          pixels_calc_ssd -> coefficient cost
// - done for a whole picture (or a band of CTU rows) in one call, with everything between the source plane and the
// results resident in HBM: one H2D copy of the source rows, kernels, D2H of costs / modes / per-TU results and the
// quantised coefficients in compact form (significance bit array + packed non-zero levels).
//
// The reference samples of every block are taken from the SOURCE picture's neighbours (open loop): the closed-loop
// dependency on the reconstruction of the neighbouring CUs is control flow of the encoder proper, which stays on the
// host; the reference arm of bench.py (oracle/ref_harness.c::refh_frame_pass) makes the same choice so both arms do
// the same arithmetic. Sample (x, y) outside the picture = the nearest picture sample (border replication).
//
// Layout. All per-block arrays are CTU-major: block index = (ctu - ctu_begin) * blocks_per_ctu + j, j = raster
// position inside the 64x64 CTU; TU j of a CTU sits at raster position j % positions and is the (j / positions)-th
// visit of that position: visit r uses the r-th best candidate mode of the rough block that covers it.
#include "common.cuh"
#include "intra.cuh"
#include <mutex>
#include <new>
#include <string.h>

namespace uvgb {

namespace {

constexpr int kCtu = 64;
constexpr int kMaxLv = UVG_B200_FRAME_MAX_LEVELS;
constexpr int kMaxStages = kMaxLv + 2;            // rough search | one per TU size | compaction

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// CTU-major block index -> top-left pixel of the block
struct Geo {
  int W, H, ctu_cols, ctu_begin;
  __device__ __forceinline__ void pos(size_t blk, int per_ctu, int n, int &X, int &Y, int &ctu, int &j) const
  {
    ctu = (int)(blk / per_ctu) + ctu_begin;
    j = (int)(blk % per_ctu);
    const int bpr = kCtu / n, p = j % (bpr * bpr);
    X = (ctu % ctu_cols) * kCtu + (p % bpr) * n;
    Y = (ctu / ctu_cols) * kCtu + (p / bpr) * n;
  }
};

// The n_best cheapest candidate modes of every block, best first; ties keep the earlier candidate (a stable sort
// of the cost list, like the insertion into the sorted mode list of search_intra.c:1064).
__global__ void __launch_bounds__(256)
frame_topk_kernel(const uint32_t *__restrict__ cost, size_t n_blocks, int k, int n_best,
                  const int8_t *__restrict__ mode_values, uint8_t *__restrict__ best, uint32_t *__restrict__ best_cost)
{
  const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  uint32_t bc[4] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu };
  int bi[4] = { 0, 0, 0, 0 };
  for (int m = 0; m < k; ++m) {
    uint32_t c = cost[b * k + m];
    int idx = m;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (s < n_best && c < bc[s]) { const uint32_t tc = bc[s]; const int ti = bi[s]; bc[s] = c; bi[s] = idx; c = tc; idx = ti; }
    }
  }
  for (int s = 0; s < n_best; ++s) {
    const int src = s < k ? s : k - 1;
    best[b * n_best + s] = (uint8_t)mode_values[bi[src]];
    best_cost[b * n_best + s] = bc[src];
  }
}

// Residual-pass operands of one TU size: the original block and the intra prediction with the TU's mode, both
// dense n*n. The reference rows come straight from the source plane (clamped), the mode from the covering rough
// block's candidate list.
struct TuSel { int src_n, src_per_ctu, n_best; const uint8_t *best; };

// shared-memory bytes of one TU of frame_tu_predict_kernel: an ODD number of 32-bit words, so that lanes working at the
// same offset of different TUs hit different banks (192 bytes for 8x8 = 48 words made it a 16-way conflict)
__host__ __device__ constexpr int tu_smem_bytes(int n)
{
  const int b = 4 * (2 * n + 12) + (n + 8) + kExtMirror + n * n;
  return ((b / 4) & 1) ? b : b + 4;
}

template <int N, int TPC>
__global__ void __launch_bounds__(256)
frame_tu_predict_kernel(const uint8_t *__restrict__ plane, int stride, Geo g, int per_ctu, size_t n_tus, TuSel sel,
                        ModeTable mt /* index = mode value 0..66 */, uint8_t *__restrict__ ref_out,
                        uint8_t *__restrict__ pred_out, uint8_t *__restrict__ mode_out)
{
  // Work item = 8 horizontally adjacent pixels of one row of a prediction in its WORK orientation (horizontal angular
  // modes are generated transposed, intra-generic.c:286-298), one dp4a per pixel (angular_row8 / planar_dc_row8 of
  // intra.cuh - the row generators of the fused rough-search kernel). Rows land in a shared-memory tile in picture
  // orientation and leave as 32-bit words.
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int n = N, RS = 2 * n + 12, HEAD = n + 8, nn = n * n;
  constexpr int TILE = 4 * RS + HEAD + kExtMirror;           // offset of a TU's prediction tile
  constexpr int per_tu = tu_smem_bytes(N);                   // [4 reference rows | HEAD | mirror | tile], odd word count
  constexpr int l2 = N == 4 ? 2 : N == 8 ? 3 : N == 16 ? 4 : 5;
  __shared__ int s_dc[TPC], s_mode[TPC], s_X[TPC], s_Y[TPC];
  constexpr int kClasses = 8;                                // planar, DC, 2 + {no PDPC, PDPC at 0 displacement, PDPC} + 3 * horizontal
  __shared__ unsigned char s_class[TPC], s_order[TPC], s_rank[TPC];
  __shared__ int s_hist[kClasses];
  const int tid = threadIdx.x;
  const size_t first = (size_t)blockIdx.x * TPC;
  if (tid >= 256 - kClasses) s_hist[255 - tid] = 0;          // (threads that have nothing to do in the next block)
  if (tid < TPC) {
    const size_t t = min(first + tid, n_tus - 1);
    int X, Y, ctu, j;
    g.pos(t, per_ctu, n, X, Y, ctu, j);
    s_X[tid] = X; s_Y[tid] = Y;
    // candidate rank = how many times this position was visited before; covering block of the source level
    const int bpr = kCtu / n, rank = min(j / (bpr * bpr), sel.n_best - 1);
    const int sb = kCtu / sel.src_n;
    const int lx = (X & (kCtu - 1)) / sel.src_n, ly = (Y & (kCtu - 1)) / sel.src_n;
    const size_t sblk = (size_t)(ctu - g.ctu_begin) * sel.src_per_ctu + ly * sb + lx;
    const int mode = sel.best[sblk * sel.n_best + rank];
    s_mode[tid] = mode;
    // code-path class of the mode: a warp's 8-pixel row items are handed out in class order (below), so that the lanes
    // of a warp run the same branch of the predictor even though every TU has its own mode
    const ModeParam mp = mt.m[mode];
    s_class[tid] = (unsigned char)(mp.kind != 2 ? mp.kind : 2 + (mp.pdpc ? (mp.sample_disp == 0 ? 1 : 2) : 0) + (mp.vertical ? 0 : 3));
  }
  __syncthreads();
  // counting sort of the CTA's TUs by class (order inside a class does not matter): a shared-memory histogram hands out
  // the ranks, the running sums over the eight classes give the bases
  if (tid < TPC) s_rank[tid] = (unsigned char)atomicAdd(&s_hist[s_class[tid]], 1);
  __syncthreads();
  if (tid < TPC) {
    const int c = s_class[tid];
    int base = 0;
#pragma unroll
    for (int k = 0; k < kClasses; ++k) base += k < c ? s_hist[k] : 0;
    s_order[base + s_rank[tid]] = (unsigned char)tid;
  }
  if (N >= 16) {
    // large TUs: few TUs per CTA, long rows - one thread per reference sample
    for (int i = tid; i < TPC * 2 * RS; i += 256) {
      const int tu = i / (2 * RS), rem = i - tu * 2 * RS, row = rem / RS, x = min(rem - row * RS, 2 * n);
      const int X = s_X[tu], Y = s_Y[tu];
      const int sx = row ? X - 1 : X - 1 + x, sy = row ? Y - 1 + x : Y - 1;
      smem_raw[tu * per_tu + row * RS + (rem - row * RS)] = __ldg(plane + (size_t)clampi(sy, 0, g.H - 1) * stride + clampi(sx, 0, g.W - 1));
    }
    __syncthreads();
    for (int i = tid; i < TPC * 2 * RS; i += 256) {          // intra_filter_reference, intra.c:190-225
      const int tu = i / (2 * RS), rem = i - tu * 2 * RS, row = rem / RS, x = rem - row * RS;
      const uint8_t *top = smem_raw + tu * per_tu, *left = top + RS, *src = row ? left : top;
      int v;
      if (x == 0) v = ((int)left[1] + 2 * (int)left[0] + (int)top[1] + 2) >> 2;
      else if (x < 2 * n) v = ((int)src[x - 1] + 2 * (int)src[x] + (int)src[x + 1] + 2) >> 2;
      else v = src[2 * n];
      smem_raw[tu * per_tu + (2 + row) * RS + x] = (uint8_t)v;
    }
  } else {
    // Reference rows from the plane, tail replicated. 1 + LP work items per TU (one pass of the CTA's 256 threads): the
    // top row (contiguous in the plane: aligned word loads + funnel shifts when the TU is away from the picture
    // borders) and LP pieces of the strided left column.
    constexpr int LP = N == 4 ? 2 : 4, Q = RS / LP;            // pieces of the left column, samples per piece
    static_assert(RS % LP == 0 && TPC * (1 + LP) <= 256, "staging items");
    for (int i = tid; i < TPC * (1 + LP); i += 256) {
      const int tu = i / (1 + LP), part = i - tu * (1 + LP);
      const int X = s_X[tu], Y = s_Y[tu];
      uint8_t *dst = smem_raw + tu * per_tu + (part ? RS : 0);
      if (part == 0 && Y >= 1 && X >= 4 && X + 2 * n + 4 <= g.W && Y - 1 < g.H) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(plane + (size_t)(Y - 1) * stride + (X - 4));   // X, stride: multiples of 4
        constexpr int NW = (2 * n + 1 + 3) / 4;                                  // words that hold top[0 .. 2n]
        uint32_t w = __ldg(src);
  #pragma unroll
        for (int j = 0; j < NW; ++j) {
          const uint32_t wn = __ldg(src + j + 1);
          uint32_t v = __funnelshift_r(w, wn, 24);                               // bytes X-1+4j .. X+2+4j
          if (j == NW - 1) v = (v & 0xffu) * 0x01010101u;                        // 2n is a multiple of 4: the last word starts with the last sample top[2n]
          *reinterpret_cast<uint32_t *>(dst + 4 * j) = v;
          w = wn;
        }
        const uint32_t last = dst[2 * n] * 0x01010101u;
  #pragma unroll
        for (int j = NW; j < RS / 4; ++j) *reinterpret_cast<uint32_t *>(dst + 4 * j) = last;
      } else if (part == 0) {
        for (int x = 0; x < RS; ++x)
          dst[x] = __ldg(plane + (size_t)clampi(Y - 1, 0, g.H - 1) * stride + clampi(X - 1 + min(x, 2 * n), 0, g.W - 1));
      } else {
        const int x0 = (part - 1) * Q;
        const uint8_t *col = plane + clampi(X - 1, 0, g.W - 1);
        if (Y >= 1 && Y - 1 + 2 * n < g.H) {                   // no clamp along the column
          const uint8_t *ps = col + (size_t)(Y - 1 + min(x0, 2 * n)) * stride;
  #pragma unroll
          for (int x = 0; x < Q; ++x) { dst[x0 + x] = __ldg(ps); if (x0 + x < 2 * n) ps += stride; }
        } else {
          for (int x = x0; x < x0 + Q; ++x) dst[x] = __ldg(col + (size_t)clampi(Y - 1 + min(x, 2 * n), 0, g.H - 1) * stride);
        }
      }
    }
    __syncthreads();
    if (N > 4) {                                               // no 4x4 mode uses the smoothed rows (intra.c:691-712: n*n > 32 / n != 4)
      for (int i = tid; i < TPC * 2; i += 256) {               // intra_filter_reference, intra.c:190-225: one thread per (TU, row)
        const int tu = i >> 1, row = i & 1;
        const uint8_t *top = smem_raw + tu * per_tu, *left = top + RS, *src = row ? left : top;
        uint8_t *dst = smem_raw + tu * per_tu + (2 + row) * RS;
        int a = src[0], bq = src[1];
        dst[0] = (uint8_t)(((int)left[1] + 2 * (int)left[0] + (int)top[1] + 2) >> 2);
        for (int x = 1; x < 2 * n; ++x) {
          const int c = src[x + 1];
          dst[x] = (uint8_t)((a + 2 * bq + c + 2) >> 2);
          a = bq; bq = c;
        }
        for (int x = 2 * n; x < RS; ++x) dst[x] = src[2 * n];
      }
    }
  }
  if (tid < TPC) {                                           // intra_pred_dc, intra.c:236-273
    const uint8_t *top = smem_raw + tid * per_tu, *left = top + RS;
    int sum = 0;
    for (int i = 1; i <= n; ++i) sum += (int)top[i] + (int)left[i];
    s_dc[tid] = (sum + n) >> (l2 + 1);
  }
  __syncthreads();
  // negative angles: main-reference indices -HEAD..-1 = the side reference projected through the inverse angle
  // (intra-generic.c:145-156)
  for (int i = tid; i < TPC * HEAD; i += 256) {
    const int tu = i / HEAD, j = i - tu * HEAD + 1;
    const ModeParam mp = mt.m[s_mode[tu]];
    if (mp.kind == 2 && mp.sample_disp < 0) {
      const uint8_t *top = smem_raw + tu * per_tu + (mp.filtered ? 2 * RS : 0), *left = top + RS;
      const uint8_t *side = mp.vertical ? left : top;
      smem_raw[tu * per_tu + 4 * RS + HEAD - j] = side[min((j * mp.inv_angle + 256) >> 9, n)];
      // angular_row8 reads the four words of a row from one base: the scratch continues with a copy of main[0 .. 12)
      if (j <= kExtMirror) smem_raw[tu * per_tu + 4 * RS + HEAD + j - 1] = (mp.vertical ? top : left)[j - 1];
    }
  }
  __syncthreads();
  constexpr int ns = n >= 8 ? n / 8 : 1, items_per_tu = n * ns;
  for (int e = tid; e < TPC * items_per_tu; e += 256) {
    const int slot = e / items_per_tu, it = e - slot * items_per_tu, r = it / ns, sx = it - r * ns;
    const int tu = s_order[slot];
    if (first + tu >= n_tus) continue;
    const ModeParam mp = mt.m[s_mode[tu]];
    const uint8_t *base = smem_raw + tu * per_tu;
    const uint8_t *top = base + (mp.filtered ? 2 * RS : 0), *left = top + RS;
    uint8_t *tile = smem_raw + tu * per_tu + TILE;
    uint32_t lo, hi;
    bool transposed = false;
    if (mp.kind == 2) {
      const uint8_t *mainr = mp.vertical ? top : left, *side = mp.vertical ? left : top;
      angular_row8(mp, mainr, side, base + 4 * RS + HEAD, r, sx, min(3 << max((int)mp.scale, 0), n), lo, hi);
      transposed = !mp.vertical;
    } else {
      planar_dc_row8(mp, top, left, s_dc[tu], n, l2, r, sx, lo, hi);
    }
    if (!transposed) {
      *reinterpret_cast<uint32_t *>(tile + r * n + 8 * sx) = lo;
      if (n >= 8) *reinterpret_cast<uint32_t *>(tile + r * n + 8 * sx + 4) = hi;
    } else {
      const int cnt = n >= 8 ? 8 : 4;
      for (int c = 0; c < cnt; ++c) tile[(8 * sx + c) * n + r] = (uint8_t)(((c < 4 ? lo : hi) >> (8 * (c & 3))) & 0xffu);
    }
  }
  __syncthreads();
  constexpr int wpt = nn / 4;                                // words per TU
  for (int e = tid; e < TPC * wpt; e += 256) {
    const int tu = e / wpt, wi = e - tu * wpt, y = (4 * wi) >> l2, x = (4 * wi) & (n - 1);
    const size_t t = first + tu;
    if (t >= n_tus) break;
    const uint8_t *tile = smem_raw + tu * per_tu + TILE;
    reinterpret_cast<uint32_t *>(pred_out + t * (size_t)nn)[wi] = *reinterpret_cast<const uint32_t *>(tile + 4 * wi);
    const int X = s_X[tu] + x, Y = s_Y[tu] + y;
    uint32_t ow;
    if (X + 3 < g.W && Y < g.H) {
      ow = __ldg(reinterpret_cast<const uint32_t *>(plane + (size_t)Y * stride + X));
    } else {
      const uint8_t *row = plane + (size_t)min(Y, g.H - 1) * stride;
      ow = (uint32_t)row[min(X, g.W - 1)] | ((uint32_t)row[min(X + 1, g.W - 1)] << 8) | ((uint32_t)row[min(X + 2, g.W - 1)] << 16) |
           ((uint32_t)row[min(X + 3, g.W - 1)] << 24);
    }
    reinterpret_cast<uint32_t *>(ref_out + t * (size_t)nn)[wi] = ow;
    if (wi == 0) mode_out[t] = (uint8_t)s_mode[tu];
  }
}

// ---- compact coefficient output -----------------------------------------------------------------------------------
// The dense coefficient arrays of the residual pass ([n_tus][n*n] int16) are mostly zeros. What crosses PCIe is
//   sig    : one bit per coefficient of the dense array (bit i of the array = bit (i & 31) of word i >> 5)
//   levels : the non-zero coefficients, TU after TU, each TU's in index order
//   off    : per TU, where its levels start
// Lane l of a warp owns coefficients [8l, 8l+8) of a 256-coefficient chunk.
template <int S /* coefficients per TU: 16, 64, 256, 1024 */>
__device__ __forceinline__ void chunk_masks(const int16_t *__restrict__ coeff, size_t chunk, size_t n_chunks, int lane,
                                            uint32_t &mask8, uint4 &v)
{
  v = make_uint4(0, 0, 0, 0);
  if (chunk < n_chunks) v = __ldg(reinterpret_cast<const uint4 *>(coeff) + chunk * 32 + lane);
  const uint32_t w[4] = { v.x, v.y, v.z, v.w };
  mask8 = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (w[i] & 0xffffu) mask8 |= 1u << (2 * i);
    if (w[i] >> 16) mask8 |= 1u << (2 * i + 1);
  }
}

// One pass over a TU size's dense coefficients: significance bits, the packed levels, and per TU where its levels start.
// A warp takes one unit (a 256-coefficient chunk; a whole TU for 32x32) per iteration and keeps its 16-byte pieces in
// registers; the CTA's eight unit totals are summed, ONE atomicAdd on the picture's level counter reserves the CTA's
// span of `levels`, and every lane stores its non-zero values at span + prefix. A TU's levels are contiguous and in
// index order; in which order the TUs (and TU sizes) lie in `levels` is whatever the atomics decide - tu_level_off[]
// says where each one starts. (The two-pass form - count kernel, device-wide scan, pack kernel - read the 111 MB of a
// 4K picture's coefficients twice and cost 95 us; this one 50.)
template <int S>
__global__ void __launch_bounds__(256)
coeff_compact_kernel(const int16_t *__restrict__ coeff, size_t n_coeffs, uint32_t *__restrict__ sig, uint32_t *__restrict__ off,
                     int16_t *__restrict__ levels, uint32_t *__restrict__ counter)
{
  __shared__ unsigned s_tot[8], s_base[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t n_chunks = (n_coeffs + 255) / 256;            // n_coeffs is a multiple of 16; a tail chunk reads zeros
  constexpr int CPT = S > 256 ? S / 256 : 1;                 // chunks per unit (32x32: the TU's 4)
  constexpr int LPT = S >= 256 ? 32 : S / 8;                 // lanes per TU inside a chunk
  const size_t n_units = (n_chunks + CPT - 1) / CPT, units_per_iter = (size_t)gridDim.x * 8;
  const size_t iters = (n_units + units_per_iter - 1) / units_per_iter;
  for (size_t it = 0; it < iters; ++it) {
    const size_t unit = it * units_per_iter + (size_t)blockIdx.x * 8 + warp, c0 = unit * CPT;
    uint4 v[CPT]; uint32_t m8[CPT]; unsigned excl[CPT], seg0[CPT], tot[CPT];
    unsigned unit_total = 0;
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc) {
      const size_t chunk = c0 + cc;
      const bool live = unit < n_units && (chunk * 256 + (size_t)lane * 8) < n_coeffs;
      m8[cc] = 0; v[cc] = make_uint4(0, 0, 0, 0);
      if (live) chunk_masks<S>(coeff, chunk, n_chunks, lane, m8[cc], v[cc]);
      uint32_t m = m8[cc] << (8 * (lane & 3));
      m |= __shfl_xor_sync(0xffffffffu, m, 1);
      m |= __shfl_xor_sync(0xffffffffu, m, 2);
      if ((lane & 3) == 0 && live) sig[chunk * 8 + (lane >> 2)] = m;
      const unsigned c = __popc(m8[cc]);
      unsigned incl = c;                                     // inclusive scan over the warp
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
      excl[cc] = incl - c;
      seg0[cc] = __shfl_sync(0xffffffffu, excl[cc], lane - (lane % LPT));
      tot[cc] = __shfl_sync(0xffffffffu, incl, 31);
      unit_total += tot[cc];
    }
    if (lane == 0) s_tot[warp] = unit_total;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned sum = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += s_tot[w];
      unsigned base = sum ? atomicAdd(counter, sum) : 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) { s_base[w] = base; base += s_tot[w]; }
    }
    __syncthreads();
    unsigned pos = s_base[warp];
#pragma unroll
    for (int cc = 0; cc < CPT; ++cc) {
      const size_t chunk = c0 + cc, first = chunk * 256 + (size_t)lane * 8;
      if (unit < n_units && first < n_coeffs) {
        if (S <= 256) { if ((lane % LPT) == 0) off[first / S] = pos + seg0[cc]; }
        else if (cc == 0 && lane == 0) off[first / S] = pos;
        if (m8[cc]) {
          int16_t *dst = levels + pos + excl[cc];
          const uint32_t w[4] = { v[cc].x, v[cc].y, v[cc].z, v[cc].w };
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (w[i] & 0xffffu) *dst++ = (int16_t)(w[i] & 0xffffu);
            if (w[i] >> 16) *dst++ = (int16_t)(w[i] >> 16);
          }
        }
      }
      pos += tot[cc];
    }
  }
}

// The packed levels are the one result whose size the host only learns after the pass. The default is a second,
// host-issued copy of exactly that size (uvg_b200_frame_wait); with UVG266_B200_FRAME_PUSH=1 a few CTAs store them
// straight into the (UVA-mapped) pinned result buffer over PCIe instead, as 16-byte posted writes, 512 contiguous bytes
// per warp instruction, the total read on the device: no host round trip, shorter latency for one picture at a time.
__global__ void __launch_bounds__(256)
levels_push_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ host_dst, const uint32_t *__restrict__ total_levels)
{
  const size_t n16 = ((size_t)*total_levels * sizeof(int16_t) + 15) / 16;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldcs(src + i);
    __stwt(host_dst + i, v);
  }
}

struct Level {
  int n = 0, k = 0, ppc = 0;
  int8_t modes[68];
  ModeTable mt;                                   // candidates in the order given
  size_t nb = 0;                                  // blocks of the whole picture
  uint8_t *d_orig = nullptr, *d_refs = nullptr, *d_best = nullptr;
  int8_t *d_modes = nullptr;
  uint32_t *d_cost = nullptr, *d_best_cost = nullptr;
  uint32_t *h_cost = nullptr, *h_best_cost = nullptr; uint8_t *h_best = nullptr;
};

struct TuSize {
  int n = 0, per_ctu = 0, src = 0;
  size_t nt = 0;
  ModeTable mt;                                   // all 67 modes of this size, index = mode value
  uint8_t *d_ref = nullptr, *d_pred = nullptr, *d_rec = nullptr, *d_mode = nullptr, *d_has = nullptr;
  int16_t *d_coeff = nullptr;
  uint32_t *d_ssd = nullptr, *d_cost = nullptr, *d_sig = nullptr;
  uint8_t *h_mode = nullptr, *h_has = nullptr;
  uint32_t *h_ssd = nullptr, *h_cost = nullptr, *h_sig = nullptr;
  int16_t *h_coeff = nullptr;                     // dense, only with UVG_B200_FRAME_DENSE_COEFF
};

}  // namespace

struct Frame {
  uvg_b200_frame_cfg cfg;
  int W = 0, H = 0, pitch = 0, ctu_cols = 0, ctu_rows = 0, n_ctu = 0;
  cudaStream_t s_in = nullptr, s_copy = nullptr;         // source rows in, results out; the kernels go into the lane
  cudaStream_t s_side[kMaxLv - 1] = {};                  // side branches of a stage (independent kernels run side by side)
  cudaEvent_t ev_stage[kMaxStages] = {}, ev_in = nullptr, ev_done = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxLv - 1] = {};
  uint8_t *d_plane = nullptr;
  Level lv[kMaxLv];
  TuSize tu[kMaxLv];
  size_t total_tus = 0, total_coeffs = 0;
  uint32_t *d_off = nullptr, *h_off = nullptr;         // per TU: where its levels start; [total TUs] = the level counter
  int16_t *d_levels = nullptr, *h_levels = nullptr;
  uvg_b200_frame_result res;
  int ctu_begin = 0, ctu_end = 0, flags = 0;
  bool in_flight = false;
  struct PassGraph {
    bool used = false, ready = false;
    int ctu_begin = 0, ctu_end = 0, flags = 0, seen = 0;
    cudaGraphExec_t kernels[kMaxStages] = {}, copies[kMaxStages] = {};
    uint64_t launches[kMaxStages] = {}, copy_launches[kMaxStages] = {};
  };
  PassGraph graphs[4];
};

static void frame_release(Frame *f)
{
  if (f->s_in) cudaStreamSynchronize(f->s_in);
  if (f->ev_t1) cudaEventSynchronize(f->ev_t1);
  if (f->s_copy) cudaStreamSynchronize(f->s_copy);
  auto D = [](void *p) { if (p) cudaFree(p); };
  auto Hh = [](void *p) { if (p) cudaFreeHost(p); };
  D(f->d_plane); D(f->d_off); D(f->d_levels);
  Hh(f->h_off); Hh(f->h_levels);
  for (int i = 0; i < kMaxLv; ++i) {
    Level &l = f->lv[i];
    D(l.d_orig); D(l.d_refs); D(l.d_best); D(l.d_cost); D(l.d_best_cost); D(l.d_modes); Hh(l.h_cost); Hh(l.h_best); Hh(l.h_best_cost);
    TuSize &t = f->tu[i];
    D(t.d_ref); D(t.d_pred); D(t.d_rec); D(t.d_mode); D(t.d_has); D(t.d_coeff); D(t.d_ssd); D(t.d_cost); D(t.d_sig);
    Hh(t.h_mode); Hh(t.h_has); Hh(t.h_ssd); Hh(t.h_cost); Hh(t.h_sig); Hh(t.h_coeff);
  }
  for (Frame::PassGraph &g : f->graphs)
    for (int s = 0; s < kMaxStages; ++s) { if (g.kernels[s]) cudaGraphExecDestroy(g.kernels[s]); if (g.copies[s]) cudaGraphExecDestroy(g.copies[s]); }
  for (cudaEvent_t e : f->ev_stage) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : { f->ev_in, f->ev_done, f->ev_t0, f->ev_t1, f->ev_fork }) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : f->ev_join) if (e) cudaEventDestroy(e);
  for (cudaStream_t q : f->s_side) if (q) cudaStreamDestroy(q);
  if (f->s_in) cudaStreamDestroy(f->s_in);
  if (f->s_copy) cudaStreamDestroy(f->s_copy);
  cudaGetLastError();
}

template <typename T> static int dalloc(T *&p, size_t count) { UVGB_TRY(cudaMalloc(reinterpret_cast<void **>(&p), count * sizeof(T) + 64)); return 0; }
template <typename T> static int halloc(T *&p, size_t count) { UVGB_TRY(cudaMallocHost(reinterpret_cast<void **>(&p), count * sizeof(T) + 64)); return 0; }

static int frame_build(Frame *f, const uvg_b200_frame_cfg *c)
{
  f->cfg = *c;
  f->W = c->width; f->H = c->height;
  f->ctu_cols = (c->width + kCtu - 1) / kCtu; f->ctu_rows = (c->height + kCtu - 1) / kCtu;
  f->n_ctu = f->ctu_cols * f->ctu_rows;
  UVGB_TRY(cudaStreamCreateWithFlags(&f->s_in, cudaStreamNonBlocking));
  for (cudaStream_t &q : f->s_side) UVGB_TRY(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
  UVGB_TRY(cudaEventCreateWithFlags(&f->ev_fork, cudaEventDisableTiming));
  for (cudaEvent_t &e : f->ev_join) UVGB_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  UVGB_TRY(cudaStreamCreateWithFlags(&f->s_copy, cudaStreamNonBlocking));
  for (cudaEvent_t &e : f->ev_stage) UVGB_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  UVGB_TRY(cudaEventCreateWithFlags(&f->ev_in, cudaEventDisableTiming));
  UVGB_TRY(cudaEventCreateWithFlags(&f->ev_done, cudaEventDisableTiming));
  UVGB_TRY(cudaEventCreate(&f->ev_t0));
  UVGB_TRY(cudaEventCreate(&f->ev_t1));
  f->pitch = (f->W + 15) & ~15;                                  // rows start 16-byte aligned: word loads of block rows
  if (dalloc(f->d_plane, (size_t)f->pitch * f->H)) return 1;
  for (int i = 0; i < c->n_levels; ++i) {
    Level &l = f->lv[i];
    l.n = c->level[i].n; l.k = c->level[i].k; l.ppc = (kCtu / l.n) * (kCtu / l.n);
    memcpy(l.modes, c->level[i].modes, sizeof(l.modes));
    l.mt.k = l.k;
    for (int m = 0; m < l.k; ++m) l.mt.m[m] = make_mode(l.n, l.modes[m]);
    l.nb = (size_t)f->n_ctu * l.ppc;
    if (dalloc(l.d_cost, l.nb * l.k) ||
        dalloc(l.d_best, l.nb * c->n_best) || dalloc(l.d_best_cost, l.nb * c->n_best) || dalloc(l.d_modes, (size_t)68) ||
        halloc(l.h_cost, l.nb * l.k) || halloc(l.h_best, l.nb * c->n_best) || halloc(l.h_best_cost, l.nb * c->n_best)) return 1;
    UVGB_TRY(cudaMemcpy(l.d_modes, l.modes, 68, cudaMemcpyHostToDevice));
  }
  for (int i = 0; i < c->n_tus; ++i) {
    TuSize &t = f->tu[i];
    t.n = c->tu[i].n; t.per_ctu = c->tu[i].per_ctu;
    // mode source: the smallest rough level that is at least as large as the TU, else the largest level
    int src = -1;
    for (int l = 0; l < c->n_levels; ++l) if (f->lv[l].n >= t.n && (src < 0 || f->lv[l].n < f->lv[src].n)) src = l;
    if (src < 0) for (int l = 0; l < c->n_levels; ++l) if (src < 0 || f->lv[l].n > f->lv[src].n) src = l;
    t.src = src;
    t.mt.k = kMaxModes;
    for (int m = 0; m < kMaxModes; ++m) t.mt.m[m] = make_mode(t.n, m);
    t.nt = (size_t)f->n_ctu * t.per_ctu;
    const size_t px = t.nt * t.n * t.n;
    if (dalloc(t.d_ref, px) || dalloc(t.d_pred, px) || dalloc(t.d_rec, px) || dalloc(t.d_mode, t.nt) || dalloc(t.d_has, t.nt) ||
        dalloc(t.d_coeff, px) || dalloc(t.d_ssd, t.nt) || dalloc(t.d_cost, t.nt) || dalloc(t.d_sig, px / 32 + 8) ||
        halloc(t.h_mode, t.nt) || halloc(t.h_has, t.nt) || halloc(t.h_ssd, t.nt) || halloc(t.h_cost, t.nt) ||
        halloc(t.h_sig, px / 32 + 8)) return 1;
    f->total_tus += t.nt; f->total_coeffs += px;
  }
  if (dalloc(f->d_off, f->total_tus + 1) || halloc(f->h_off, f->total_tus + 1) ||
      dalloc(f->d_levels, f->total_coeffs + 8) || halloc(f->h_levels, f->total_coeffs + 8)) return 1;
  return 0;
}

static int cfg_check(const uvg_b200_frame_cfg *c)
{
  if (!c) { set_error("frame: NULL configuration"); return 1; }
  if (c->bitdepth != 8) { set_error("frame: the pipeline is built for 8-bit video (bitdepth %d)", c->bitdepth); return 1; }
  if (c->width < 16 || c->height < 16 || c->width > 16384 || c->height > 16384) { set_error("frame: %dx%d out of range", c->width, c->height); return 1; }
  if (c->n_levels < 1 || c->n_levels > kMaxLv || c->n_tus < 0 || c->n_tus > kMaxLv) { set_error("frame: 1..4 rough levels, 0..4 TU sizes"); return 1; }
  if (c->n_best < 1 || c->n_best > 4) { set_error("frame: n_best must be 1..4"); return 1; }
  for (int i = 0; i < c->n_levels; ++i) {
    const int n = c->level[i].n, k = c->level[i].k;
    if (n != 8 && n != 16 && n != 32) { set_error("frame: rough level size %d unsupported (8, 16, 32)", n); return 1; }
    if (k < 1 || k > 67) { set_error("frame: 1..67 candidate modes per level"); return 1; }
    for (int m = 0; m < k; ++m) if (c->level[i].modes[m] < 0 || c->level[i].modes[m] > 66) { set_error("frame: mode out of range"); return 1; }
  }
  for (int i = 0; i < c->n_tus; ++i) {
    const int n = c->tu[i].n;
    if (n != 4 && n != 8 && n != 16 && n != 32) { set_error("frame: TU size %d unsupported", n); return 1; }
    if (c->tu[i].per_ctu < 1 || c->tu[i].per_ctu > 4 * (kCtu / n) * (kCtu / n)) { set_error("frame: per_ctu out of range"); return 1; }
  }
  return 0;
}

template <int S>
static void launch_compact(const int16_t *coeff, size_t n_coeffs, uint32_t *sig, uint32_t *off, int16_t *levels, uint32_t *counter,
                           cudaStream_t st)
{
  constexpr int CPT = S > 256 ? S / 256 : 1;
  const size_t units = ((n_coeffs + 255) / 256 + CPT - 1) / CPT;
  size_t grid = (units + 7) / 8;
  const size_t cap = (size_t)sm_count() * 8;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  coeff_compact_kernel<S><<<(unsigned)grid, 256, 0, st>>>(coeff, n_coeffs, sig, off, levels, counter);
}

static int compact(int n, const int16_t *coeff, size_t n_coeffs, uint32_t *sig, uint32_t *off, int16_t *levels, uint32_t *counter,
                   cudaStream_t st)
{
  switch (n) {
    case 4: launch_compact<16>(coeff, n_coeffs, sig, off, levels, counter, st); break;
    case 8: launch_compact<64>(coeff, n_coeffs, sig, off, levels, counter, st); break;
    case 16: launch_compact<256>(coeff, n_coeffs, sig, off, levels, counter, st); break;
    default: launch_compact<1024>(coeff, n_coeffs, sig, off, levels, counter, st); break;
  }
  UVGB_LAUNCH_CHECK();
  return 0;
}

}  // namespace uvgb

using namespace uvgb;

struct uvg_b200_frame { Frame f; };

extern "C" uvg_b200_frame *uvg_b200_frame_create(const uvg_b200_frame_cfg *cfg)
{
  if (ensure_init()) return nullptr;
  if (cfg_check(cfg)) return nullptr;
  uvg_b200_frame *h = new (std::nothrow) uvg_b200_frame();
  if (!h) { set_error("frame: out of memory"); return nullptr; }
  if (frame_build(&h->f, cfg)) { frame_release(&h->f); delete h; return nullptr; }
  return h;
}

extern "C" void uvg_b200_frame_destroy(uvg_b200_frame *h)
{
  if (!h) return;
  frame_release(&h->f);
  delete h;
}

// ---- how a pass is scheduled -----------------------------------------------------------------------------------------
// Every frame object's KERNELS go into one process-wide compute stream, in submission order (the lane). Two pictures'
// kernels never share the SMs: with a stream per picture the older picture's chain of short dependent kernels queues
// behind the younger one's long rough-search waves, both finish together, and the copies at either end of a pass (source
// rows in, results out: ~0.45 ms per 4K picture) find no kernels to hide under. In the lane picture B's kernels start
// the moment A's end, B's source rows arrive (s_in) while A computes, and A's results leave (s_copy) while B computes.
// A pass is cut into stages - rough search | one per TU size | compaction - and after each stage an event releases that
// stage's D2H copies on the picture's own s_copy. Kernels and copies of a stage are each captured into a CUDA graph the
// second time a (band, flags) configuration is seen (the first run performs the lazy one-time initialisation of the
// kernels it calls) and replayed from then on.
struct ComputeLane { cudaStream_t st = nullptr; std::mutex mu; };
static ComputeLane *lane()
{
  static ComputeLane l;
  static std::once_flag once;
  std::call_once(once, [] { if (cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking) != cudaSuccess) l.st = nullptr; });
  return l.st ? &l : nullptr;
}

constexpr int kPushCtas = 64;
static bool levels_pushed()
{
  // opt-in: measured on the B200 with two pictures in flight the push kernel's CTAs, resident beside the next picture's
  // kernels for the ~0.3 ms the PCIe writes take, slow those by ~20 % (1.15 ms per 4K picture against 0.98 with the
  // host-issued copy, which costs the picture one round trip of latency but leaves the SMs alone)
  static const bool on = getenv("UVG266_B200_FRAME_PUSH") && !strcmp(getenv("UVG266_B200_FRAME_PUSH"), "1");
  return on;
}

static int n_stages(const Frame *f) { return 1 + f->cfg.n_tus; }

static size_t tus_before(const Frame *f, int band_ctus, int i)
{
  size_t b = 0;
  for (int j = 0; j < i; ++j) b += (size_t)band_ctus * f->tu[j].per_ctu;
  return b;
}

// kernels of stage s, into `st`
static int stage_kernels(Frame *f, int s, int ctu_begin, int ctu_end, cudaStream_t st)
{
  const uvg_b200_frame_cfg &c = f->cfg;
  const int W = f->W, H = f->H, band_ctus = ctu_end - ctu_begin;
  if (s == 0) {
    // ---- rough search: the levels are independent of each other, so level i > 0 runs on a side branch (fork / join by
    // events; inside a capture these become parallel branches of the stage's graph): one level's last, partly filled wave
    // of CTAs is topped up by the other level's instead of leaving SMs idle
    static const bool serial = getenv("UVG266_B200_FRAME_SIDE") && !strcmp(getenv("UVG266_B200_FRAME_SIDE"), "0");
    if (c.n_levels > 1 && !serial) UVGB_TRY(cudaEventRecord(f->ev_fork, st));
    for (int i = 0; i < c.n_levels; ++i) {
      Level &l = f->lv[i];
      const size_t nb = (size_t)band_ctus * l.ppc;
      cudaStream_t q = (i == 0 || serial) ? st : f->s_side[i - 1];
      if (q != st) UVGB_TRY(cudaStreamWaitEvent(q, f->ev_fork, 0));
      const PlaneSrc ps{ f->d_plane, f->pitch, W, H, f->ctu_cols, ctu_begin };
      if (launch_intra_rough_plane(l.n, ps, l.modes, l.k, l.d_cost, nb, q)) return 1;
      frame_topk_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, q>>>(l.d_cost, nb, l.k, c.n_best, l.d_modes, l.d_best, l.d_best_cost);
      UVGB_LAUNCH_CHECK();
      if (q != st) {
        UVGB_TRY(cudaEventRecord(f->ev_join[i - 1], q));
        UVGB_TRY(cudaStreamWaitEvent(st, f->ev_join[i - 1], 0));
      }
    }
    return 0;
  }
  if (s <= c.n_tus) {                                      // ---- residual pass of one TU size
    const int i = s - 1;
    TuSize &t = f->tu[i];
    const Level &sl = f->lv[t.src];
    const Geo g{ W, H, f->ctu_cols, ctu_begin };
    const size_t nt = (size_t)band_ctus * t.per_ctu;
    // TUs per CTA chosen so that the 8-pixel row items of a CTA's TUs fill its 256 threads
    const int nn = t.n * t.n, tpc = t.n == 4 ? 64 : t.n == 8 ? 32 : t.n == 16 ? 8 : 2;
    const size_t smem = (size_t)tpc * tu_smem_bytes(t.n);
    const TuSel sel{ sl.n, sl.ppc, c.n_best, sl.d_best };
    const unsigned grid = (unsigned)((nt + tpc - 1) / tpc);
    if (t.n == 4) frame_tu_predict_kernel<4, 64><<<grid, 256, smem, st>>>(f->d_plane, f->pitch, g, t.per_ctu, nt, sel, t.mt, t.d_ref, t.d_pred, t.d_mode);
    else if (t.n == 8) frame_tu_predict_kernel<8, 32><<<grid, 256, smem, st>>>(f->d_plane, f->pitch, g, t.per_ctu, nt, sel, t.mt, t.d_ref, t.d_pred, t.d_mode);
    else if (t.n == 16) frame_tu_predict_kernel<16, 8><<<grid, 256, smem, st>>>(f->d_plane, f->pitch, g, t.per_ctu, nt, sel, t.mt, t.d_ref, t.d_pred, t.d_mode);
    else frame_tu_predict_kernel<32, 2><<<grid, 256, smem, st>>>(f->d_plane, f->pitch, g, t.per_ctu, nt, sel, t.mt, t.d_ref, t.d_pred, t.d_mode);
    UVGB_LAUNCH_CHECK();
    uvg_b200_blocks R, P, O;
    R.base = t.d_ref; P.base = t.d_pred; O.base = t.d_rec;
    for (uvg_b200_blocks *b : { &R, &P, &O }) { b->stride = (uint32_t)t.n; b->blocks_per_row = 1; b->col_pitch = 0; b->row_pitch = (uint64_t)nn; }
    if (uvg_b200_quantize_residual_dev(&c.quant, 1 /* CU_INTRA */, 0, t.n, t.n, 0, 0, 0, &R, &P, &O, t.d_coeff, t.d_has, t.d_ssd,
                                       t.d_cost, c.cost_weights, nt, st)) return 1;
    // the picture's level counter lives behind the offsets: d_off[total TUs of the band]
    uint32_t *counter = f->d_off + tus_before(f, band_ctus, c.n_tus);
    if (i == 0) UVGB_TRY(cudaMemsetAsync(counter, 0, sizeof(uint32_t), st));
    return compact(t.n, t.d_coeff, nt * (size_t)nn, t.d_sig, f->d_off + tus_before(f, band_ctus, i), f->d_levels, counter, st);
  }
  return 0;
}

// D2H copies that stage s releases, into `sc`; *any tells whether there were any
static int stage_copies(Frame *f, int s, int ctu_begin, int ctu_end, int flags, cudaStream_t sc, bool *any)
{
  const uvg_b200_frame_cfg &c = f->cfg;
  const int band_ctus = ctu_end - ctu_begin;
  *any = false;
  if (flags & UVG_B200_FRAME_NO_D2H) return 0;
  *any = true;
  if (s == 0) {
    for (int i = 0; i < c.n_levels; ++i) {
      Level &l = f->lv[i];
      const size_t nb = (size_t)band_ctus * l.ppc;
      // what search_intra_rough hands back is the list of best candidates with their costs (search_intra.c:1215-1223);
      // the cost of every (block, mode) only crosses PCIe on request
      if (flags & UVG_B200_FRAME_ALL_COSTS) UVGB_TRY(cudaMemcpyAsync(l.h_cost, l.d_cost, nb * l.k * sizeof(uint32_t), cudaMemcpyDeviceToHost, sc));
      UVGB_TRY(cudaMemcpyAsync(l.h_best, l.d_best, nb * c.n_best, cudaMemcpyDeviceToHost, sc));
      UVGB_TRY(cudaMemcpyAsync(l.h_best_cost, l.d_best_cost, nb * c.n_best * sizeof(uint32_t), cudaMemcpyDeviceToHost, sc));
    }
    return 0;
  }
  if (s <= c.n_tus) {
    TuSize &t = f->tu[s - 1];
    const size_t nt = (size_t)band_ctus * t.per_ctu, nn = (size_t)t.n * t.n;
    UVGB_TRY(cudaMemcpyAsync(t.h_mode, t.d_mode, nt, cudaMemcpyDeviceToHost, sc));
    UVGB_TRY(cudaMemcpyAsync(t.h_has, t.d_has, nt, cudaMemcpyDeviceToHost, sc));
    UVGB_TRY(cudaMemcpyAsync(t.h_ssd, t.d_ssd, nt * 4, cudaMemcpyDeviceToHost, sc));
    UVGB_TRY(cudaMemcpyAsync(t.h_cost, t.d_cost, nt * 4, cudaMemcpyDeviceToHost, sc));
    UVGB_TRY(cudaMemcpyAsync(t.h_sig, t.d_sig, (nt * nn + 31) / 32 * 4, cudaMemcpyDeviceToHost, sc));
    if (flags & UVG_B200_FRAME_DENSE_COEFF) UVGB_TRY(cudaMemcpyAsync(t.h_coeff, t.d_coeff, nt * nn * 2, cudaMemcpyDeviceToHost, sc));
    const size_t before = tus_before(f, band_ctus, s - 1);
    UVGB_TRY(cudaMemcpyAsync(f->h_off + before, f->d_off + before, nt * sizeof(uint32_t), cudaMemcpyDeviceToHost, sc));
    if (s == c.n_tus) {                                      // the last size: the level counter is final
      const size_t tu_total = tus_before(f, band_ctus, c.n_tus);
      UVGB_TRY(cudaMemcpyAsync(f->h_off + tu_total, f->d_off + tu_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, sc));
      if (levels_pushed()) {
        levels_push_kernel<<<kPushCtas, 256, 0, sc>>>(reinterpret_cast<const uint4 *>(f->d_levels), reinterpret_cast<uint4 *>(f->h_levels),
                                                      f->d_off + tu_total);
        UVGB_LAUNCH_CHECK();
      }
    }
    return 0;
  }
  return 0;
}

// run `body` into `st`: directly, or (capture = true) captured into *exec and launched
template <typename Body>
static int run_or_capture(cudaStream_t st, bool capture, cudaGraphExec_t *exec, Body body)
{
  if (!capture) return body();
  cudaGraph_t graph = nullptr;
  UVGB_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  const int rc = body();
  const cudaError_t ce = cudaStreamEndCapture(st, &graph);
  if (rc || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    if (!rc) set_error("frame: graph capture failed: %s", cudaGetErrorString(ce));
    cudaGetLastError();
    return 1;
  }
  const cudaError_t ie = cudaGraphInstantiate(exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) { *exec = nullptr; set_error("frame: graph instantiation failed: %s", cudaGetErrorString(ie)); return 1; }
  UVGB_TRY(cudaGraphLaunch(*exec, st));
  return 0;
}

extern "C" int uvg_b200_frame_submit(uvg_b200_frame *h, const void *src, size_t src_stride, int ctu_row_begin,
                                     int ctu_row_end, int flags)
{
  if (!h) { set_error("frame: NULL handle"); return 1; }
  if (ensure_init()) return 1;                           // also selects the library's device for a new calling thread
  Frame *f = &h->f;
  ComputeLane *ln = lane();
  if (!ln) { set_error("frame: no compute stream"); return 1; }
  if (f->in_flight) { set_error("frame: submit while a pass is in flight (call uvg_b200_frame_wait first)"); return 1; }
  if (ctu_row_end <= 0 || ctu_row_end > f->ctu_rows) ctu_row_end = f->ctu_rows;
  if (ctu_row_begin < 0 || ctu_row_begin >= ctu_row_end) { set_error("frame: CTU row range [%d,%d) is empty", ctu_row_begin, ctu_row_end); return 1; }
  const int W = f->W, H = f->H;
  const int ctu_begin = ctu_row_begin * f->ctu_cols, ctu_end = ctu_row_end * f->ctu_cols;
  f->ctu_begin = ctu_begin; f->ctu_end = ctu_end; f->flags = flags;
  if ((flags & UVG_B200_FRAME_DENSE_COEFF) && !(flags & UVG_B200_FRAME_NO_D2H))
    for (int i = 0; i < f->cfg.n_tus; ++i) {
      TuSize &t = f->tu[i];
      if (!t.h_coeff && halloc(t.h_coeff, t.nt * (size_t)t.n * t.n)) return 1;
    }
  if (!(flags & UVG_B200_FRAME_NO_H2D)) {
    if (!src) { set_error("frame: NULL source plane"); return 1; }
    // rows the band reads: one above, and below it what the left reference column of a 32x32 block at the band's
    // last rows reaches (Y - 1 + 2n)
    const int y0 = max(0, ctu_row_begin * kCtu - 1), y1 = min(H, ctu_row_end * kCtu + kCtu);
    UVGB_TRY(cudaMemcpy2DAsync(f->d_plane + (size_t)y0 * f->pitch, f->pitch, static_cast<const uint8_t *>(src) + (size_t)y0 * src_stride,
                               src_stride, W, y1 - y0, cudaMemcpyHostToDevice, f->s_in));
  }
  UVGB_TRY(cudaEventRecord(f->ev_in, f->s_in));
  const int key_flags = flags & ~UVG_B200_FRAME_NO_H2D;
  Frame::PassGraph *pg = nullptr;
  for (Frame::PassGraph &g : f->graphs)
    if (g.used && g.ctu_begin == ctu_begin && g.ctu_end == ctu_end && g.flags == key_flags) pg = &g;
  if (!pg)
    for (Frame::PassGraph &g : f->graphs)
      if (!g.used) { g = Frame::PassGraph(); g.used = true; g.ctu_begin = ctu_begin; g.ctu_end = ctu_end; g.flags = key_flags; pg = &g; break; }
  static const bool no_graph = getenv("UVG266_B200_FRAME_GRAPH") && !strcmp(getenv("UVG266_B200_FRAME_GRAPH"), "0");
  const bool replay = pg && pg->ready, capture = pg && !pg->ready && pg->seen >= 1 && !no_graph;
  const int ns = n_stages(f);
  {
    std::lock_guard<std::mutex> lock(ln->mu);              // one picture's kernels at a time, in submission order
    UVGB_TRY(cudaStreamWaitEvent(ln->st, f->ev_in, 0));
    UVGB_TRY(cudaEventRecord(f->ev_t0, ln->st));
    for (int s = 0; s < ns; ++s) {
      if (replay) {
        UVGB_TRY(cudaGraphLaunch(pg->kernels[s], ln->st));
        g_launches.fetch_add(pg->launches[s], std::memory_order_relaxed);
      } else {
        const uint64_t l0 = g_launches.load();
        if (run_or_capture(ln->st, capture, capture ? &pg->kernels[s] : nullptr, [&] { return stage_kernels(f, s, ctu_begin, ctu_end, ln->st); })) return 1;
        if (capture) pg->launches[s] = g_launches.load() - l0;
      }
      UVGB_TRY(cudaEventRecord(f->ev_stage[s], ln->st));
    }
    UVGB_TRY(cudaEventRecord(f->ev_t1, ln->st));
  }
  for (int s = 0; s < ns; ++s) {
    UVGB_TRY(cudaStreamWaitEvent(f->s_copy, f->ev_stage[s], 0));
    if (replay) {
      if (pg->copies[s]) { UVGB_TRY(cudaGraphLaunch(pg->copies[s], f->s_copy)); g_launches.fetch_add(pg->copy_launches[s], std::memory_order_relaxed); }
    } else {
      bool any = false;
      const uint64_t l0 = g_launches.load();
      // (a stage without copies is not captured: probe first)
      if (flags & UVG_B200_FRAME_NO_D2H) continue;
      if (run_or_capture(f->s_copy, capture, capture ? &pg->copies[s] : nullptr, [&] { return stage_copies(f, s, ctu_begin, ctu_end, flags, f->s_copy, &any); })) return 1;
      if (capture) pg->copy_launches[s] = g_launches.load() - l0;
    }
  }
  UVGB_TRY(cudaEventRecord(f->ev_done, f->s_copy));
  if (pg) { if (capture) pg->ready = true; else if (!replay) pg->seen++; }
  f->in_flight = true;
  return 0;
}

extern "C" int uvg_b200_frame_wait(uvg_b200_frame *h, uvg_b200_frame_result *out)
{
  if (!h) { set_error("frame: NULL handle"); return 1; }
  if (ensure_init()) return 1;
  Frame *f = &h->f;
  if (!f->in_flight) { set_error("frame: wait without submit"); return 1; }
  f->in_flight = false;
  const uvg_b200_frame_cfg &c = f->cfg;
  const bool d2h = !(f->flags & UVG_B200_FRAME_NO_D2H);
  const int band_ctus = f->ctu_end - f->ctu_begin;
  uvg_b200_frame_result &r = f->res;
  memset(&r, 0, sizeof(r));
  r.ctu_begin = f->ctu_begin; r.ctu_end = f->ctu_end;
  const size_t tu_total = tus_before(f, band_ctus, c.n_tus);
  UVGB_TRY(cudaEventSynchronize(f->ev_done));             // kernels of the pass and the copies they released
  if (d2h && c.n_tus) {
    const uint64_t total = f->h_off[tu_total];
    r.n_levels = total;
    if (!levels_pushed()) {
      // without the push kernel the host, which only now knows how many levels there are, issues one more copy of
      // exactly that size
      if (total) UVGB_TRY(cudaMemcpyAsync(f->h_levels, f->d_levels, total * sizeof(int16_t), cudaMemcpyDeviceToHost, f->s_copy));
      UVGB_TRY(cudaStreamSynchronize(f->s_copy));
    }
  }
  float ms = 0.f;
  UVGB_TRY(cudaEventSynchronize(f->ev_t1));               // (recorded just after the last stage's event)
  UVGB_TRY(cudaEventElapsedTime(&ms, f->ev_t0, f->ev_t1));  // the pass's kernels, first to last
  r.device_ms = ms;
  if (d2h) {
    size_t base = 0;
    for (int i = 0; i < c.n_levels; ++i) {
      r.rough_cost[i] = (f->flags & UVG_B200_FRAME_ALL_COSTS) ? f->lv[i].h_cost : nullptr;
      r.rough_best[i] = f->lv[i].h_best; r.rough_best_cost[i] = f->lv[i].h_best_cost;
    }
    for (int i = 0; i < c.n_tus; ++i) {
      TuSize &t = f->tu[i];
      r.tu_mode[i] = t.h_mode; r.tu_has_coeffs[i] = t.h_has; r.tu_ssd[i] = t.h_ssd; r.tu_cost[i] = t.h_cost; r.tu_sig[i] = t.h_sig;
      r.tu_level_off[i] = f->h_off + base;
      r.tu_coeff_dense[i] = (f->flags & UVG_B200_FRAME_DENSE_COEFF) ? t.h_coeff : nullptr;
      base += (size_t)band_ctus * t.per_ctu;
    }
    r.levels = f->h_levels;
  }
  if (out) *out = r;
  return 0;
}

extern "C" int uvg_b200_frame_device_buffers(uvg_b200_frame *h, uvg_b200_frame_device *out)
{
  if (!h || !out) { set_error("frame: NULL argument"); return 1; }
  Frame *f = &h->f;
  memset(out, 0, sizeof(*out));
  out->plane = f->d_plane;
  for (int i = 0; i < f->cfg.n_levels; ++i) { out->rough_cost[i] = f->lv[i].d_cost; out->rough_best[i] = f->lv[i].d_best; }
  for (int i = 0; i < f->cfg.n_tus; ++i) {
    out->tu_pred[i] = f->tu[i].d_pred; out->tu_rec[i] = f->tu[i].d_rec; out->tu_coeff[i] = f->tu[i].d_coeff;
  }
  out->stream = lane() ? lane()->st : nullptr;           // the stream every pass's kernels run in
  return 0;
}

extern "C" int uvg_b200_device_read(void *host_dst, const void *dev_src, size_t bytes)
{
  if (ensure_init()) return 1;
  if (!bytes) return 0;
  if (!host_dst || !dev_src) { set_error("device_read: NULL argument"); return 1; }
  UVGB_TRY(cudaMemcpy(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost));
  return 0;
}

// Host-side inverse of the compact coefficient format: dense[i] for the n_coeffs coefficients of one TU size.
extern "C" void uvg_b200_frame_expand_coeffs(const uint32_t *sig, const uint32_t *level_off, const int16_t *levels,
                                             int coeffs_per_tu, size_t n_tus, int16_t *dense)
{
  for (size_t t = 0; t < n_tus; ++t) {
    const int16_t *lv = levels + level_off[t];
    for (int i = 0; i < coeffs_per_tu; ++i) {
      const size_t gidx = t * (size_t)coeffs_per_tu + i;
      const int bit = (sig[gidx >> 5] >> (gidx & 31)) & 1;
      dense[gidx] = bit ? *lv++ : 0;
    }
  }
}
