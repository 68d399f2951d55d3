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
#include <string.h>
#include <stdlib.h>

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

// ---- TMA (cp.async.bulk) staged variant -------------------------------------------------------
// Same computation, but the 2 KB candidate chunks are brought into shared memory by the TMA unit
// (1-D bulk copy, UBLKCP in SASS) into a 3-deep per-warp ring guarded by mbarriers, instead of
// LDG -> registers -> STS: no load/store instructions and no staging registers are spent on the
// candidate stream, and three chunks are in flight per warp.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int kStages = 3;

template <int N, bool ALIGNED>
__global__ void __launch_bounds__(kWarps * 32)
rough_bulk_kernel(Blocks A, const uint8_t *__restrict__ cand, uint32_t *__restrict__ out, uint32_t n_items, FastDiv kd)
{
  constexpr int IPW = kChunkBytes / (N * N);
  __shared__ __align__(128) uint8_t smem[kWarps][kStages][kChunkBytes];
  __shared__ __align__(8) uint64_t bars[kWarps][kStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_chunks = (n_items + IPW - 1) / IPW;
  const size_t total_bytes = (size_t)n_items * (size_t)(N * N);
  const uint32_t stride_chunks = gridDim.x * kWarps;
  const uint32_t first = blockIdx.x * kWarps + warp;

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&bars[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncwarp();

  auto issue = [&](uint32_t chunk, int stage) {      // lane 0 only
    const size_t off = (size_t)chunk * kChunkBytes;
    size_t bytes = total_bytes - off;
    if (bytes > kChunkBytes) bytes = kChunkBytes;     // the tail chunk is shorter (multiple of 64 bytes)
    mbar_expect_tx(&bars[warp][stage], (uint32_t)bytes);
    bulk_g2s(smem[warp][stage], cand + off, (uint32_t)bytes, &bars[warp][stage]);
  };

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      const uint64_t c = (uint64_t)first + (uint64_t)s * stride_chunks;
      if (c < n_chunks) issue((uint32_t)c, s);
    }
  }
  uint32_t it = 0;
  for (uint32_t ch = first; ch < n_chunks; ch += stride_chunks, ++it) {
    const int stage = it % kStages;
    const uint32_t parity = (it / kStages) & 1;
    mbar_wait(&bars[warp][stage], parity);
    const uint8_t *slice = smem[warp][stage];

    uint32_t wo[8][2], wc[8][2];
    if constexpr (N == 8) {
      const uint32_t item = ch * IPW + lane;
      const bool valid = item < n_items;
      const int m = (lane >> 1) & 3;
      const uint8_t *pa = static_cast<const uint8_t *>(A.base) + A.offset(valid ? fdiv(item, kd) : 0);
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        const int piece = qd ^ m;
        const uint4 v = *reinterpret_cast<const uint4 *>(slice + lane * 64 + piece * 16);
        wc[2 * qd][0] = v.x; wc[2 * qd][1] = v.y; wc[2 * qd + 1][0] = v.z; wc[2 * qd + 1][1] = v.w;
        load_orig_row8<8, ALIGNED>(pa + (size_t)(2 * piece) * A.stride, wo[2 * qd][0], wo[2 * qd][1]);
        load_orig_row8<8, ALIGNED>(pa + (size_t)(2 * piece + 1) * A.stride, wo[2 * qd + 1][0], wo[2 * qd + 1][1]);
      }
      // the slice is in registers now: hand the buffer back to the TMA for chunk it + kStages
      __syncwarp();
      if (lane == 0) {
        const uint64_t nc = (uint64_t)ch + (uint64_t)kStages * stride_chunks;
        if (nc < n_chunks) issue((uint32_t)nc, stage);
      }
      unsigned sad;
      const unsigned satd = satd8_words_u8<true>(wo, wc, sad);
      if (valid) out[item] = min(satd, 2 * sad);
    } else {
      const int itx = lane >> 2, bx = lane & 1, by = (lane >> 1) & 1;
      const uint32_t item = ch * IPW + itx;
      const bool valid = item < n_items;
      const int f = ((itx & 3) << 1) | by;
      const uint8_t *pa = static_cast<const uint8_t *>(A.base) + A.offset(valid ? fdiv(item, kd) : 0) +
                          (size_t)(8 * by) * A.stride + 8 * bx;
      const uint8_t *pc = slice + itx * 256 + by * 128 + bx * 8;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int row = r ^ f;
        const uint2 v = *reinterpret_cast<const uint2 *>(pc + row * 16);
        wc[r][0] = v.x; wc[r][1] = v.y;
        load_orig_row8<16, ALIGNED>(pa + (size_t)row * A.stride, wo[r][0], wo[r][1]);
      }
      __syncwarp();
      if (lane == 0) {
        const uint64_t nc = (uint64_t)ch + (uint64_t)kStages * stride_chunks;
        if (nc < n_chunks) issue((uint32_t)nc, stage);
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
  // Default: LDG -> registers -> STS staging. UVG266_B200_ROUGH=tma selects the cp.async.bulk
  // variant for A/B runs: measured on B200 it is 2-7 % slower (8x8: 0.0811 vs 0.0753 ms, 16x16:
  // 0.0825 vs 0.0808 ms per launch) - the kernel is issue-bound on the Hadamard, not on the
  // staging, and the mbarrier round trip per 2 KB chunk costs more than the 8 LDG/STS it removes.
  static const char *mode = getenv("UVG266_B200_ROUGH");
  const bool use_bulk = mode && !strcmp(mode, "tma");
  const unsigned g = (unsigned)grid, b = kWarps * 32;
  if (use_bulk) {
    if (n == 8) {
      if (aligned) rough_bulk_kernel<8, true><<<g, b, 0, st>>>(A, cb, out, ni, kd);
      else rough_bulk_kernel<8, false><<<g, b, 0, st>>>(A, cb, out, ni, kd);
    } else {
      if (aligned) rough_bulk_kernel<16, true><<<g, b, 0, st>>>(A, cb, out, ni, kd);
      else rough_bulk_kernel<16, false><<<g, b, 0, st>>>(A, cb, out, ni, kd);
    }
  } else if (n == 8) {
    if (aligned) rough_staged_kernel<8, true><<<g, b, 0, st>>>(A, cb, out, ni, kd);
    else rough_staged_kernel<8, false><<<g, b, 0, st>>>(A, cb, out, ni, kd);
  } else {
    if (aligned) rough_staged_kernel<16, true><<<g, b, 0, st>>>(A, cb, out, ni, kd);
    else rough_staged_kernel<16, false><<<g, b, 0, st>>>(A, cb, out, ni, kd);
  }
  UVGB_LAUNCH_CHECK();
  return 0;
}

}  // namespace uvgb
