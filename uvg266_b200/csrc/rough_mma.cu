// rough_mma.cu - rough-search cost min(SATD, 2*SAD) with the Hadamard on the INT8 tensor cores.
//
// Semantics: search_intra.c:149-163 (get_cost) over uvg_satd_NxN / uvg_sad_NxN duals
// (picture-generic.c:483-603): per 8x8 sub-block  (sum |H8 D H8^T| + 2) >> 2, summed; SAD summed.
//
// Why tensor cores here: the packed-ALU Hadamard of rough.cu costs ~860 issue slots per 8x8
// candidate and the kernel sits at 100 % issue utilisation at 40-43 % of the HBM roofline - the
// only way up is fewer instructions per byte. The 2-D Hadamard of an 8x8 block is one product
// with the 64x64 +-1 matrix H8 (x) H8 (entry (k,n) = (-1)^popcount(k & n)), and by linearity
// H(O - P)H^T = H O H^T - H P H^T, so the 8-bit pixels themselves are the u8 operand:
//   mma.sync.m16n8k32.u8.s8: 16 blocks (rows) x 64 pixels (K) x 64 coefficients (N). K is split
//   into even / odd pixel rows (the lane's two 16-byte vectors): S0, S1 for the four even
//   coefficient columns are 8 IMMAs, and the odd columns differ only in the sign of the odd pixel
//   rows, so coefficient 2j = S0 + S1 and 2j+1 = S0 - S1 (one accumulator butterfly instead of
//   8 more IMMAs). 13 -> 8 IMMAs per tile left the kernel time unchanged (0.0906 / 0.0720 ms):
//   the kernel is issue / latency bound (55 % issue, L1 56 %), not tensor-pipe bound;
//   row 0 of every tile = the original block, rows 1..15 = 15 candidates of that block;
//   epilogue per lane: 32 x VABSDIFF-accumulate |Z - T_O| (T_O broadcast from row 0 by shuffle),
//   8 x VABSDIFF4 for the SAD on the same fragment registers, one packed quad reduction.
// ~6 issue slots per candidate instead of ~860; no shared memory; every prediction byte is read
// from HBM exactly once, as a 4-byte fragment word straight into the MMA operand register.
// Any ordering / sign convention of the Hadamard basis gives the same sum of magnitudes, and
// sum|c| <= 64 * ||D||_2 <= 130560 < 2^17, which is what lets SATD and SAD share one register
// through the reduction.
#include "common.cuh"
#include <cstdlib>

namespace uvgb {

__device__ __forceinline__ void imma32(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kRowsPerTile = 15;      // candidates per 16-row tile (row 0 = the original)

struct TileSrc {                      // where this lane's two tile rows live
  const uint8_t *lo, *hi;             // row g, row g + 8
  uint32_t lo_stride;                 // orig stride for lane group 0, N otherwise
};

// Lane (g, t) owns pixel rows 2t and 2t+1 of every 8x8 sub-block of tile rows g and g+8, so its
// fragment words are whole 16-byte vectors in memory:
//   N = 8 : the two pixel rows are contiguous            -> one LDG.128 per tile row;
//   N >= 16: one LDG.128 per pixel row covers the sub-blocks sx = 2p and 2p+1, handled as a pair.
// K index of a fragment byte: pixel 16t + 4q + i of the sub-block with q = 2s + h
// (k-step s = pixel-row parity, half h = left / right 4 pixels).
template <int N>
struct Unit {
  static constexpr int PAIR = N >= 16 ? 2 : 1;
  uint4 lo[PAIR], hi[PAIR];           // [pixel-row parity] for N >= 16; [0] = both rows for N = 8
};

template <int N>
__global__ void __launch_bounds__(128)
rough_mma_kernel(Blocks A, const uint8_t *__restrict__ cand, uint32_t *__restrict__ out, uint32_t n_blocks, int k)
{
  constexpr int NS = N / 8, PAIR = Unit<N>::PAIR, UNITS = NS * NS / PAIR, UPR = NS / PAIR;   // units per tile / per sub-block row
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // B operand = column n = 8j + g of H8 (x) H8 at rows m = 16t + 8s + 4h + i: sign = parity of
  // popcount(m & n) = (i & g & 3) ^ (h & g>>2) ^ (s & j) ^ (t & j>>1): byte pattern and the (h, j>>1)
  // dependent flips are per-lane registers; the (s & j & 1) flip is a butterfly on the accumulators.
  uint32_t pat = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) pat |= ((__popc(i & g & 3) & 1) ? 0xffu : 0x01u) << (8 * i);
  uint32_t wp[2][4];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int flip = ((h & (g >> 2)) ^ __popc(t & jj)) & 1;
      wp[h][jj] = flip ? pat ^ 0xfefefefeu : pat;
    }
  const int tiles = (k + kRowsPerTile - 1) / kRowsPerTile;
  const uint32_t warp_global = blockIdx.x * 4 + (threadIdx.x >> 5), n_warps = gridDim.x * 4;

  auto source = [&](uint32_t b, int tile, TileSrc &src) {
    const uint8_t *pc = cand + (size_t)b * k * (N * N);
    const int c_lo = min(tile * kRowsPerTile + g - 1, k - 1), c_hi = min(tile * kRowsPerTile + g + 7, k - 1);
    src.hi = pc + (size_t)c_hi * (N * N);
    if (g == 0) { src.lo = static_cast<const uint8_t *>(A.base) + A.offset(b); src.lo_stride = A.stride; }
    else { src.lo = pc + (size_t)c_lo * (N * N); src.lo_stride = N; }
  };
  auto load = [&](const TileSrc &src, int u, Unit<N> &r) {
    if (N == 8) {
      r.hi[0] = __ldg(reinterpret_cast<const uint4 *>(src.hi + 16 * t));
      if (g == 0) {                                       // the original block is strided
        const uint2 r0 = __ldg(reinterpret_cast<const uint2 *>(src.lo + (size_t)(2 * t) * src.lo_stride));
        const uint2 r1 = __ldg(reinterpret_cast<const uint2 *>(src.lo + (size_t)(2 * t + 1) * src.lo_stride));
        r.lo[0] = make_uint4(r0.x, r0.y, r1.x, r1.y);
      } else {
        r.lo[0] = __ldg(reinterpret_cast<const uint4 *>(src.lo + 16 * t));
      }
    } else {
      const int sy = u / UPR, sxp = u - sy * UPR;
#pragma unroll
      for (int s = 0; s < PAIR; ++s) {
        const int y = sy * 8 + 2 * t + s, x = 16 * sxp;
        r.lo[s] = __ldg(reinterpret_cast<const uint4 *>(src.lo + (size_t)y * src.lo_stride + x));
        r.hi[s] = __ldg(reinterpret_cast<const uint4 *>(src.hi + y * N + x));
      }
    }
  };
  // fragment words of sub-block p of a unit: a[s] = {lo h0, hi h0, lo h1, hi h1}
  auto frag = [&](const Unit<N> &r, int p, uint32_t (&a)[2][4]) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const uint4 l = N == 8 ? r.lo[0] : r.lo[s], h = N == 8 ? r.hi[0] : r.hi[s];
      const int w = N == 8 ? 2 * s : 2 * p;               // first word of (row parity s, sub-block p)
      a[s][0] = w == 0 ? l.x : l.z; a[s][2] = w == 0 ? l.y : l.w;
      a[s][1] = w == 0 ? h.x : h.z; a[s][3] = w == 0 ? h.y : h.w;
    }
  };

  uint32_t b = warp_global;
  if (b >= n_blocks) return;
  int tile = 0, u = 0;
  TileSrc src;
  source(b, 0, src);
  Unit<N> cur, nxt;
  load(src, 0, cur);
  unsigned tot_lo = 0, tot_hi = 0, sad_lo = 0, sad_hi = 0;
  while (true) {
    // advance the (block, tile, unit) cursor and prefetch the next unit's vectors
    uint32_t nb_ = b;
    int ntile = tile, nu = u + 1;
    if (nu == UNITS) { nu = 0; if (++ntile == tiles) { ntile = 0; nb_ = b + n_warps; } }
    const bool more = nb_ < n_blocks;
    TileSrc nsrc = src;
    if (more) {
      if (nu == 0) source(nb_, ntile, nsrc);
      load(nsrc, nu, nxt);
    }
#pragma unroll
    for (int p = 0; p < PAIR; ++p) {
      uint32_t a[2][4];
      frag(cur, p, a);
      // k-step s holds the pixel rows of parity s, so column tile j (vertical frequency v = j) sees
      // (-1)^(s & j): the even rows contribute the same to tiles 2jj and 2jj+1, the odd rows with
      // opposite sign. Eight IMMAs give S0 (even rows) and S1 (odd rows) of the four even tiles; one
      // butterfly on the accumulators gives all eight tiles: Z[2jj] = S0 + S1, Z[2jj+1] = S0 - S1.
      int acc[8][4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        int s0[4] = { 0, 0, 0, 0 }, s1[4] = { 0, 0, 0, 0 };
        imma32(s0, a[0], wp[0][jj], wp[1][jj]);
        imma32(s1, a[1], wp[0][jj], wp[1][jj]);
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc[2 * jj][c] = s0[c] + s1[c]; acc[2 * jj + 1][c] = s0[c] - s1[c]; }
      }
      unsigned sl = 0, sh = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int t0 = __shfl_sync(0xffffffffu, acc[j][0], t), t1 = __shfl_sync(0xffffffffu, acc[j][1], t);
        sl = __sad(acc[j][0], t0, sl); sl = __sad(acc[j][1], t1, sl);
        sh = __sad(acc[j][2], t0, sh); sh = __sad(acc[j][3], t1, sh);
        if (j == 0 && t == 0) {                            // DC (n = 0) counts a quarter (picture-generic.c:341-343)
          const unsigned dl_ = abs(acc[0][0] - t0), dh_ = abs(acc[0][2] - t0);
          sl = sl - dl_ + (dl_ >> 2); sh = sh - dh_ + (dh_ >> 2);
        }
      }
      unsigned dl = 0, dh = 0;
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t o = __shfl_sync(0xffffffffu, a[s][2 * h], t);
          dl = __vsadu4(a[s][2 * h], o) + dl;
          dh = __vsadu4(a[s][2 * h + 1], o) + dh;
        }
      unsigned pl = sl | (dl << 17), ph = sh | (dh << 17);
      pl += __shfl_xor_sync(0xffffffffu, pl, 1); ph += __shfl_xor_sync(0xffffffffu, ph, 1);
      pl += __shfl_xor_sync(0xffffffffu, pl, 2); ph += __shfl_xor_sync(0xffffffffu, ph, 2);
      tot_lo += ((pl & 0x1ffffu) + 2) >> 2; sad_lo += pl >> 17;
      tot_hi += ((ph & 0x1ffffu) + 2) >> 2; sad_hi += ph >> 17;
    }
    if (u == UNITS - 1) {
      if (t == 0) {
        const int c_lo = tile * kRowsPerTile + g - 1, c_hi = tile * kRowsPerTile + g + 7;
        uint32_t *o = out + (size_t)b * k;
        if (g != 0 && c_lo < k) o[c_lo] = min(tot_lo, 2 * sad_lo);
        if (c_hi < k) o[c_hi] = min(tot_hi, 2 * sad_hi);
      }
      tot_lo = tot_hi = sad_lo = sad_hi = 0;
    }
    if (!more) break;
    b = nb_; tile = ntile; u = nu; src = nsrc; cur = nxt;
  }
}

// Returns -1 when the layout does not qualify (caller falls through to the staged ALU kernel).
int launch_rough_mma(const Blocks &A, const Blocks &B, uint32_t *out, size_t n_items, int k, int n, cudaStream_t st)
{
  if (n != 8 && n != 16 && n != 32) return -1;
  if (B.stride != (uint32_t)n || B.bpr != 1 || B.row_pitch != (uint64_t)(n * n)) return -1;   // dense candidates
  const unsigned am = n == 8 ? 7 : 15;                    // vector width of the original block's rows
  if (reinterpret_cast<uintptr_t>(B.base) & 15) return -1;
  if ((reinterpret_cast<uintptr_t>(A.base) & am) || (A.stride & am) || (A.col_pitch & am) || (A.row_pitch & am)) return -1;
  const size_t n_blocks = n_items / (size_t)k;
  if (n_blocks * (size_t)k != n_items || n_blocks >= (1ull << 31)) return -1;
  size_t grid = (n_blocks + 3) / 4;
  static const int cap_mult = getenv("UVG266_B200_MMA_CAP") ? atoi(getenv("UVG266_B200_MMA_CAP")) : 24;
  const size_t cap = (size_t)sm_count() * (size_t)(cap_mult > 0 ? cap_mult : 24);
  if (grid > cap) grid = cap;
  const uint8_t *cb = static_cast<const uint8_t *>(B.base);
  if (n == 8) rough_mma_kernel<8><<<(unsigned)grid, 128, 0, st>>>(A, cb, out, (uint32_t)n_blocks, k);
  else if (n == 16) rough_mma_kernel<16><<<(unsigned)grid, 128, 0, st>>>(A, cb, out, (uint32_t)n_blocks, k);
  else rough_mma_kernel<32><<<(unsigned)grid, 128, 0, st>>>(A, cb, out, (uint32_t)n_blocks, k);
  UVGB_LAUNCH_CHECK();
  return 0;
}

}  // namespace uvgb
