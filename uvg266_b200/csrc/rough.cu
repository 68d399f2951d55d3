// rough.cu - intra rough-search cost min(SATD, 2*SAD) for 8-bit 8x8 / 16x16 blocks with the
// candidate predictions staged through shared memory (get_cost_dual, search_intra.c:133-192,
// on satd_NxN_dual picture-generic.c:371-410 and sad_NxN_dual :1091-1113).
//
// Why staging: candidates are dense n*n blocks, item i at byte i*n*n. A thread that owns a whole
// 8x8 block (or one 8x8 sub-block of a 16x16) reads 8-byte row pieces 64 (or 16) bytes apart;
// issued straight to global memory every request touches 32 different 32-byte sectors for 256
// useful bytes, and ncu showed L1 moving 8x the algorithmic bytes (l1tex 74% busy, DRAM 19%).
// Here each warp instead pulls its next 2 KB of candidates with four fully coalesced 16-byte
// loads per lane (prefetched one chunk ahead into registers), parks them in its private 2 KB
// shared-memory slice, and the owners read their rows from there.
//
// Bank conflicts are avoided without padding (so the slice stays a linear image of global
// memory) by reading rows in a dyadically shifted order, which leaves SATD and SAD unchanged
// (hadamard.cuh): for 8x8, lane b reads 16-byte piece q^((b>>1)&3); for 16x16, lane
// (item i, sub-block bx,by) reads row r^(((i&3)<<1)|by).
#include "common.cuh"
#include "hadamard.cuh"

namespace uvgb {

constexpr int kChunkBytes = 2048;     // per warp and iteration
constexpr int kWarps = 4;

// item / k for 32-bit item counts: floor(item / k) = (item * magic) >> 40 for all item < 2^31 / k ... (host checks)
struct FastDiv { uint32_t k; uint64_t magic; };
__device__ __forceinline__ uint32_t fdiv(uint32_t x, const FastDiv &d)
{
  return d.k == 1 ? x : (uint32_t)__umul64hi((uint64_t)x, d.magic);
}

template <int ROWS, bool ALIGNED>
__device__ __forceinline__ void load_orig_row8(const uint8_t *row, uint32_t &w0, uint32_t &w1)
{
  if (ALIGNED) {
    const uint2 o = __ldg(reinterpret_cast<const uint2 *>(row));
    w0 = o.x; w1 = o.y;
  } else {
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { lo |= (uint32_t)__ldg(row + j) << (8 * j); hi |= (uint32_t)__ldg(row + 4 + j) << (8 * j); }
    w0 = lo; w1 = hi;
  }
}

template <int N, bool ALIGNED>
__global__ void __launch_bounds__(kWarps * 32)
rough_staged_kernel(Blocks A, const uint8_t *__restrict__ cand, uint32_t *__restrict__ out, uint32_t n_items, FastDiv kd)
{
  constexpr int IPW = kChunkBytes / (N * N);          // items per warp-chunk: 32 (8x8) or 8 (16x16)
  __shared__ __align__(16) uint8_t smem[kWarps][kChunkBytes];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t *slice = smem[warp];
  const uint32_t n_chunks = (n_items + IPW - 1) / IPW;
  const size_t total_bytes = (size_t)n_items * (size_t)(N * N);
  const uint32_t stride_chunks = gridDim.x * kWarps;
  uint32_t ch = blockIdx.x * kWarps + warp;

  auto fetch = [&](uint32_t chunk, uint4 (&pf)[4]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const size_t off = (size_t)chunk * kChunkBytes + (size_t)c * 512 + (size_t)lane * 16;
      pf[c] = off < total_bytes ? __ldg(reinterpret_cast<const uint4 *>(cand + off)) : make_uint4(0, 0, 0, 0);
    }
  };

  uint4 pf[4];
  if (ch < n_chunks) fetch(ch, pf);
  for (; ch < n_chunks; ch += stride_chunks) {
    __syncwarp();                                      // previous iteration's readers are done
#pragma unroll
    for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4 *>(slice + c * 512 + lane * 16) = pf[c];
    __syncwarp();
    if (ch + stride_chunks < n_chunks) fetch(ch + stride_chunks, pf);   // overlap with the compute below

    uint32_t wo[8][2], wc[8][2];
    if constexpr (N == 8) {
      const uint32_t item = ch * IPW + lane;
      const bool valid = item < n_items;
      const int m = (lane >> 1) & 3;
      const uint8_t *pa = static_cast<const uint8_t *>(A.base) + A.offset(valid ? fdiv(item, kd) : 0);
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        const int piece = qd ^ m;                      // rows 2*piece, 2*piece+1
        const uint4 v = *reinterpret_cast<const uint4 *>(slice + lane * 64 + piece * 16);
        wc[2 * qd][0] = v.x; wc[2 * qd][1] = v.y; wc[2 * qd + 1][0] = v.z; wc[2 * qd + 1][1] = v.w;
        load_orig_row8<8, ALIGNED>(pa + (size_t)(2 * piece) * A.stride, wo[2 * qd][0], wo[2 * qd][1]);
        load_orig_row8<8, ALIGNED>(pa + (size_t)(2 * piece + 1) * A.stride, wo[2 * qd + 1][0], wo[2 * qd + 1][1]);
      }
      unsigned sad;
      const unsigned satd = satd8_words_u8<true>(wo, wc, sad);
      if (valid) out[item] = min(satd, 2 * sad);
    } else {
      const int it = lane >> 2, bx = lane & 1, by = (lane >> 1) & 1;
      const uint32_t item = ch * IPW + it;
      const bool valid = item < n_items;
      const int f = ((it & 3) << 1) | by;
      const uint8_t *pa = static_cast<const uint8_t *>(A.base) + A.offset(valid ? fdiv(item, kd) : 0) +
                          (size_t)(8 * by) * A.stride + 8 * bx;
      const uint8_t *pc = slice + it * 256 + by * 128 + bx * 8;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int row = r ^ f;
        const uint2 v = *reinterpret_cast<const uint2 *>(pc + row * 16);
        wc[r][0] = v.x; wc[r][1] = v.y;
        load_orig_row8<16, ALIGNED>(pa + (size_t)row * A.stride, wo[r][0], wo[r][1]);
      }
      unsigned sad;
      unsigned satd = satd8_words_u8<true>(wo, wc, sad);
      satd = group_sum<4>(satd); sad = group_sum<4>(sad);
      if (valid && (lane & 3) == 0) out[item] = min(satd, 2 * sad);
    }
  }
}

// Returns -1 when the layout is not the dense / aligned one this kernel is written for.
int launch_rough_staged(const Blocks &A, const Blocks &B, uint32_t *out, size_t n_items, int k, int n,
                        cudaStream_t st)
{
  if (n != 8 && n != 16) return -1;
  if (B.stride != (uint32_t)n || B.bpr != 1 || B.row_pitch != (uint64_t)(n * n)) return -1;
  if (reinterpret_cast<uintptr_t>(B.base) & 15) return -1;
  if (n_items >= (1ull << 31)) return -1;               // 32-bit item arithmetic in the kernel
  const int ipw = kChunkBytes / (n * n);
  const size_t n_chunks = (n_items + ipw - 1) / ipw;
  size_t grid = (n_chunks + kWarps - 1) / kWarps;
  const size_t cap = (size_t)sm_count() * 12;          // resident CTAs; warps loop over chunks
  if (grid > cap) grid = cap;
  FastDiv kd;
  kd.k = (uint32_t)k;
  kd.magic = k > 1 ? (~0ull / (uint64_t)k) + 1 : 0;     // exact for x < 2^32 (ceil(2^64 / k))
  const bool aligned = (reinterpret_cast<uintptr_t>(A.base) & 7) == 0 && (A.stride & 7) == 0 &&
                       (A.col_pitch & 7) == 0 && (A.row_pitch & 7) == 0;
  const uint8_t *cb = static_cast<const uint8_t *>(B.base);
  const uint32_t ni = (uint32_t)n_items;
  if (n == 8) {
    if (aligned) rough_staged_kernel<8, true><<<(unsigned)grid, kWarps * 32, 0, st>>>(A, cb, out, ni, kd);
    else rough_staged_kernel<8, false><<<(unsigned)grid, kWarps * 32, 0, st>>>(A, cb, out, ni, kd);
  } else {
    if (aligned) rough_staged_kernel<16, true><<<(unsigned)grid, kWarps * 32, 0, st>>>(A, cb, out, ni, kd);
    else rough_staged_kernel<16, false><<<(unsigned)grid, kWarps * 32, 0, st>>>(A, cb, out, ni, kd);
  }
  UVGB_LAUNCH_CHECK();
  return 0;
}

}  // namespace uvgb
