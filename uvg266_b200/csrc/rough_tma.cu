// rough_tma.cu - rough-search cost min(SATD, 2*SAD) on the 5th-generation tensor cores, fed by tensor-map TMA.
//
// Semantics: search_intra.c:133-160 (get_cost_dual) over the satd / sad NxN duals (picture-generic.c:256-348 8x8
// Hadamard incl. the |dc| >> 2 rule, :371-410, :1091-1113): per 8x8 sub-block (sum |H8 D H8^T| + 2) >> 2, summed.
//
// The 2-D Hadamard of an 8x8 block is ONE product with the 64x64 +-1 matrix H8 (x) H8 (entry (p, c) =
// (-1)^popcount(p & c)), and by linearity H(O - P)H^T = H O H^T - H P H^T: the 8-bit pixels are the u8 operand exactly as
// they lie in memory, the originals' transform is a table.
//
// Tile = 128 rows x 128 bytes, K-major, SWIZZLE_128B (the canonical UMMA operand layout = what the TMA writes):
//   rows   0..111  candidate bytes, 128 contiguous bytes per row:
//                    N = 8 : two consecutive 8x8 candidates            (224 candidates per tile)
//                    N = 16: eight pixel rows of a 16x16 candidate = its sub-blocks (sy, 0) and (sy, 1)  (56 candidates)
//   rows 112..127  the ORIGINAL blocks the tile's candidates belong to, in the same row format (16 rows suffice for
//                  k >= 10 candidates per block) - so H O H^T comes out of the same MMA as rows 112..127 of D.
//   B (128 x 128 s8, shared memory, built once per CTA): block-diagonal, column n = 64 s + c takes the 64 pixels of
//   half s of the row (N = 8: bytes 64 s ..; N = 16: columns 8 s .. 8 s + 7 of the 8 pixel rows) with sign
//   (-1)^popcount(p & c): D[row][64 s + c] = coefficient c of the row's sub-block s.
// One tcgen05.mma.cta_group::1.kind::i8 tile M128 N128 K128 (4 instructions of K = 32) per 14 KB of candidates.
//
// Warp-specialised persistent CTA, mbarriers only in the steady state:
//   warp 8, one lane : TMA producer - two cp.async.bulk.tensor.2d (UTMALDG) per tile: a 128 B x 112 row box of the
//                      candidate matrix and a 128 B x 16 row box of the originals, both complete_tx on the stage's
//                      full barrier; 4-stage ring. No thread touches global memory for operands.
//   warp 9, one lane : MMA issuer - 4 x UTCIMMA into one of two 128-column TMEM accumulators, tcgen05.commit.
//   warps 0-7        : two epilogue groups (alternate tiles), thread = tile row. Rows 112..127 publish H O H^T of
//                      their blocks to shared memory (one named barrier per tile and group); every candidate row then
//                      reads its 128 coefficients from TMEM (tcgen05.ld), accumulates |D - T_O| against its block's
//                      table row with the |dc| >> 2 rule, takes the SAD from its own swizzled row and its original's
//                      row of the same tile (conflict-free LDS.128), and stores min(SATD, 2 SAD).
// Nothing but the TMA reads HBM; the candidate bytes are read exactly once.
#include "common.cuh"
#include <cuda.h>
#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace uvgb {

namespace {

__device__ __forceinline__ uint32_t s32addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bar_init(uint64_t *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(s32addr(bar)), "r"(count));
}
__device__ __forceinline__ void bar_arrive(uint64_t *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(s32addr(bar)) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(s32addr(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a descriptor or tensor-map mistake must end in a trap, not in a hung GPU.
__device__ __forceinline__ void bar_wait(uint64_t *bar, uint32_t parity)
{
  uint32_t ok = 0;
  for (int spin = 0; spin < (1 << 16) && !ok; ++spin) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.b32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(s32addr(bar)), "r"(parity), "r"(20000u) : "memory");
  }
  if (!ok) __trap();
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, int cols)
{
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(s32addr(dst_smem)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, int cols)
{
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(taddr), "r"(cols));
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n"
               :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(s32addr(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16])
{
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// 2-D tiled TMA load (UTMALDG): box of the tensor map at (c0 bytes, c1 rows) -> shared memory, completes on `bar`.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
               :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(s32addr(bar)) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start address >> 4
// [0,14), LBO >> 4 [16,30) (unused: K fits one swizzle atom), SBO >> 4 [32,46) = 8 rows x 128 B, version 1 [46,48),
// layout type SWIZZLE_128B = 2 [61,64).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr)
{
  return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (2 << 4), A = u8 (0 << 7), B = s8 (1 << 10), both
// K-major, N = 128 (N >> 3 << 17), M = 128 (M >> 4 << 24).
constexpr uint32_t kIdesc = (2u << 4) | (0u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

constexpr int kRows = 128, kRowBytes = 128, kTileBytes = kRows * kRowBytes, kMaxOrigRows = 16,
              kStages = 4, kAccCols = 128, kTmemCols = 256, kThreads = 320, kCtasPerSm = 2;
constexpr int kSmemB = kRows * kRowBytes;                                  // 16 KB
constexpr int kTOHalf = 68, kTOPitch = 2 * kTOHalf;                      // ints per 64-coefficient half / per table row;                                              // ints per table row (padded: rows of a warp's stores / loads spread over the banks)
constexpr int kSmemTO = 2 * kMaxOrigRows * kTOPitch * 4;                   // two groups x (up to) 16 rows
constexpr int kSmemTotal = kSmemB + kStages * kTileBytes + kSmemTO;

__device__ __forceinline__ int sw128(int row, int chunk) { return row * kRowBytes + ((chunk ^ (row & 7)) << 4); }

struct FastDivK { uint32_t k; uint64_t magic; };      // floor(x / k) for x < 2^32
__device__ __forceinline__ uint32_t fdivk(uint32_t x, const FastDivK &d)
{
  return d.k == 1 ? x : (uint32_t)__umul64hi((uint64_t)x, d.magic);
}

}  // namespace

// n_cands: candidates in total (n_blocks * k). Candidate c of the launch belongs to block c / k.
// OR = rows of the tile given to the originals: 8 (k >= 20 candidates per block: 120 candidate rows per tile) or 16 (k >= 10).
template <int N, int OR>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
rough_tma_kernel(const __grid_constant__ CUtensorMap map_cand, const __grid_constant__ CUtensorMap map_orig,
                 uint32_t *__restrict__ out, uint32_t n_cands, FastDivK kd)
{
  constexpr int kCandRows = kRows - OR;
  constexpr int CPR = N == 8 ? 2 : 1;             // candidates per tile row (N = 8) ...
  constexpr int RPC = N == 8 ? 1 : 2;             // ... or tile rows per candidate (N = 16)
  constexpr int CAND_PER_TILE = kCandRows * CPR / RPC;
  // SWIZZLE_128B operands need 1024-byte alignment; declared like this the compiler also keeps the shared address space
  // of everything derived from it (LDS / STS instead of generic loads)
  extern __shared__ __align__(1024) uint8_t dyn[];
  uint8_t *const sB = dyn;
  uint8_t *const sA = dyn + kSmemB;
  int *const sTO = reinterpret_cast<int *>(dyn + kSmemB + kStages * kTileBytes);     // [group][16][kTOPitch]
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], acc_full[2], acc_empty[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // B[n][kk], n = 64 s + c: the 64 pixels of half s of a row, sign (-1)^popcount(p & c), zero for the other half
  for (int i = tid; i < 128 * 8; i += kThreads) {
    const int n = i >> 3, ch = i & 7, s = n >> 6, c = n & 63;
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t v = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = 16 * ch + 4 * q + j;
        int half, p;
        if (N == 8) { half = kk >> 6; p = kk & 63; }
        else { const int y = kk >> 4, x = kk & 15; half = x >> 3; p = y * 8 + (x & 7); }
        const uint32_t b = half != s ? 0u : ((__popc(p & c) & 1) ? 0xffu : 0x01u);
        v |= b << (8 * j);
      }
      w[q] = v;
    }
    *reinterpret_cast<uint4 *>(sB + sw128(n, ch)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (tid == 0) {
    if (s32addr(dyn) & 1023) __trap();
    for (int i = 0; i < kStages; ++i) { bar_init(&full_bar[i], 1); bar_init(&empty_bar[i], 128); }
    for (int i = 0; i < 2; ++i) { bar_init(&acc_full[i], 1); bar_init(&acc_empty[i], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncwarp();
  if (warp == 9) tmem_alloc(&s_tmem, kTmemCols);
  fence_async_smem();                                  // B was written through the generic proxy, the MMA reads it through the async proxy
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = s_tmem;
  const uint32_t n_tiles = (n_cands + CAND_PER_TILE - 1) / CAND_PER_TILE;
  const uint32_t tile0 = blockIdx.x, tstep = gridDim.x;
  const uint32_t my_tiles = tile0 < n_tiles ? (n_tiles - tile0 + tstep - 1) / tstep : 0;

  if (warp < 8) {
    // ===== epilogue: thread = tile row; two groups of 128 threads take alternate tiles =====
    const int r = tid & 127, grp = tid >> 7, wq = warp & 3;
    const uint32_t tlane = tmem + ((uint32_t)(wq * 32) << 16) + grp * kAccCols;
    for (uint32_t i = grp; i < my_tiles; i += 2) {
      const uint32_t tile = tile0 + i * tstep, st = i % kStages;
      const uint8_t *tileA = sA + st * kTileBytes;
      // One table per group and ONE barrier per tile: the table of tile i + 2 is written only after acc_full(i + 2), i.e.
      // after every thread of the group has arrived on acc_empty(i) - which each does after its last table read of tile i.
      int *const myTO = sTO + grp * (kMaxOrigRows * kTOPitch);
      const uint32_t c_first = tile * CAND_PER_TILE, b_first = fdivk(c_first, kd);
      // first row of the originals' box in units of original-matrix rows (the producer computes the same)
      const uint32_t orow0 = N == 8 ? (b_first >> 1) : 2 * b_first;
      // which original-matrix row and half holds the original of each of the row's two 64-coefficient halves
      uint32_t cand0, torow[2], tohalf[2];
      bool live[2];
      if (N == 8) {
        cand0 = c_first + 2 * r;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t c = cand0 + h;
          live[h] = r < kCandRows && c < n_cands;
          const uint32_t b = fdivk(live[h] ? c : c_first, kd);
          torow[h] = (b >> 1) - orow0; tohalf[h] = b & 1;
        }
      } else {
        cand0 = c_first + (r >> 1);
        const bool lv = r < kCandRows && cand0 < n_cands;
        const uint32_t b = fdivk(lv ? cand0 : c_first, kd);
        live[0] = live[1] = lv;
        torow[0] = torow[1] = 2 * b + (r & 1) - orow0; tohalf[0] = 0; tohalf[1] = 1;
      }
      // ---- SAD first: it only needs the stage (the TMA's bytes), not the MMA -----------------------------------
      bar_wait(&full_bar[st], (i / kStages) & 1);
      unsigned sad[2] = { 0, 0 };
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int orow = kCandRows + (int)torow[h];
        unsigned sd[4] = { 0, 0, 0, 0 };
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 p = *reinterpret_cast<const uint4 *>(tileA + sw128(r, 4 * h + c));
          const uint4 o = *reinterpret_cast<const uint4 *>(tileA + sw128(orow, 4 * (int)tohalf[h] + c));
          sd[0] = __vsadu4(p.x, o.x) + sd[0]; sd[1] = __vsadu4(p.y, o.y) + sd[1];
          sd[2] = __vsadu4(p.z, o.z) + sd[2]; sd[3] = __vsadu4(p.w, o.w) + sd[3];
        }
        sad[h] = live[h] ? (sd[0] + sd[1]) + (sd[2] + sd[3]) : 0;
      }
      bar_wait(&acc_full[i & 1], (i >> 1) & 1);
      fence_after_sync();
      bar_arrive(&empty_bar[st]);                              // stage read by this thread and consumed by the MMA: the TMA may refill it
      // ---- rows 112..127: H O H^T of the tile's blocks -> shared memory --------------------------------
      // TMEM loads are software-pipelined: tcgen05.wait::ld covers every outstanding load, so the next chunk is
      // issued right AFTER the wait for the current one and flies while the current one is consumed.
      if (wq == 3) {
        int z[2][16];
        tmem_ld16(tlane, z[0]);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          tmem_ld_wait();
          if (q < 7) tmem_ld16(tlane + 16 * (q + 1), z[(q + 1) & 1]);
          if (r >= kCandRows) {
            int4 *dst = reinterpret_cast<int4 *>(myTO + (r - kCandRows) * kTOPitch + (q >> 2) * kTOHalf + 16 * (q & 3));
            const int(&v)[16] = z[q & 1];
            dst[0] = make_int4(v[0], v[1], v[2], v[3]); dst[1] = make_int4(v[4], v[5], v[6], v[7]);
            dst[2] = make_int4(v[8], v[9], v[10], v[11]); dst[3] = make_int4(v[12], v[13], v[14], v[15]);
          }
        }
      }
      // the first chunk of this row's own coefficients is on its way while the group waits for the table
      int z[2][16];
      tmem_ld16(tlane, z[0]);
      asm volatile("bar.sync %0, 128;\n" :: "r"(1 + grp) : "memory");
      // ---- candidate rows: sum |D - T_O| -------------------------------------------------------------------------
      unsigned satd[2] = { 0, 0 };
      const int *T[2] = { myTO + torow[0] * kTOPitch + tohalf[0] * kTOHalf, myTO + torow[1] * kTOPitch + tohalf[1] * kTOHalf };
      unsigned ac[4] = { 0, 0, 0, 0 }, dc = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int h = q >> 2;
        const int4 t0 = *reinterpret_cast<const int4 *>(T[h] + 16 * (q & 3)), t1 = *reinterpret_cast<const int4 *>(T[h] + 16 * (q & 3) + 4),
                   t2 = *reinterpret_cast<const int4 *>(T[h] + 16 * (q & 3) + 8), t3 = *reinterpret_cast<const int4 *>(T[h] + 16 * (q & 3) + 12);
        tmem_ld_wait();
        if (q < 7) {
          tmem_ld16(tlane + 16 * (q + 1), z[(q + 1) & 1]);
        } else {
          fence_before_sync();
          bar_arrive(&acc_empty[i & 1]);                      // accumulator drained: the next MMA into it may start
        }
        const int(&v)[16] = z[q & 1];
        if ((q & 3) == 0) dc = (unsigned)abs(v[0] - t0.x);
        ac[0] = __sad(v[0], t0.x, ac[0]); ac[1] = __sad(v[1], t0.y, ac[1]); ac[2] = __sad(v[2], t0.z, ac[2]); ac[3] = __sad(v[3], t0.w, ac[3]);
        ac[0] = __sad(v[4], t1.x, ac[0]); ac[1] = __sad(v[5], t1.y, ac[1]); ac[2] = __sad(v[6], t1.z, ac[2]); ac[3] = __sad(v[7], t1.w, ac[3]);
        ac[0] = __sad(v[8], t2.x, ac[0]); ac[1] = __sad(v[9], t2.y, ac[1]); ac[2] = __sad(v[10], t2.z, ac[2]); ac[3] = __sad(v[11], t2.w, ac[3]);
        ac[0] = __sad(v[12], t3.x, ac[0]); ac[1] = __sad(v[13], t3.y, ac[1]); ac[2] = __sad(v[14], t3.z, ac[2]); ac[3] = __sad(v[15], t3.w, ac[3]);
        if ((q & 3) == 3) {
          const unsigned acc = (ac[0] + ac[1]) + (ac[2] + ac[3]);
          satd[h] = live[h] ? (acc - dc + (dc >> 2) + 2) >> 2 : 0;                // picture-generic.c:341-345
          ac[0] = ac[1] = ac[2] = ac[3] = 0;
        }
      }
      if (N == 8) {
        const uint32_t c0 = min(satd[0], 2 * sad[0]), c1 = min(satd[1], 2 * sad[1]);
        if (live[1]) *reinterpret_cast<uint2 *>(out + cand0) = make_uint2(c0, c1);       // cand0 is even: 8-byte aligned
        else if (live[0]) out[cand0] = c0;
      } else {
        unsigned s2 = satd[0] + satd[1], d2 = sad[0] + sad[1];
        s2 += __shfl_xor_sync(0xffffffffu, s2, 1); d2 += __shfl_xor_sync(0xffffffffu, d2, 1);
        if (live[0] && (r & 1) == 0) out[cand0] = min(s2, 2 * d2);
      }
    }
  } else if (warp == 8) {
    // ===== TMA producer (one lane) =====
    if (lane == 0) {
      for (uint32_t i = 0; i < my_tiles; ++i) {
        const uint32_t st = i % kStages, tile = tile0 + i * tstep;
        if (i >= (uint32_t)kStages) bar_wait(&empty_bar[st], ((i / kStages) - 1) & 1);
        const uint32_t c_first = tile * CAND_PER_TILE, b_first = fdivk(c_first, kd);
        const uint32_t orow0 = N == 8 ? (b_first >> 1) : 2 * b_first;
        const uint32_t dst = s32addr(sA + st * kTileBytes);
        bar_expect_tx(&full_bar[st], kTileBytes);
        tma_load_2d(dst, &map_cand, 0, (int)(tile * kCandRows), &full_bar[st]);
        tma_load_2d(dst + kCandRows * kRowBytes, &map_orig, 0, (int)orow0, &full_bar[st]);
      }
    }
  } else if (lane == 0) {
    // ===== MMA issuer (warp 9, one lane) =====
    const uint64_t bd = smem_desc_sw128(s32addr(sB));
    for (uint32_t i = 0; i < my_tiles; ++i) {
      const uint32_t st = i % kStages;
      bar_wait(&full_bar[st], (i / kStages) & 1);
      if (i >= 2) bar_wait(&acc_empty[i & 1], ((i >> 1) - 1) & 1);
      fence_after_sync();
      const uint64_t ad = smem_desc_sw128(s32addr(sA + st * kTileBytes));
      const uint32_t d = tmem + (i & 1) * kAccCols;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_i8(d, ad + 2 * ks, bd + 2 * ks, kIdesc, ks > 0);   // 32 bytes of K per MMA
      umma_commit(&acc_full[i & 1]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 9) tmem_free(tmem, kTmemCols);
}

// ---- 8x8 blocks, BLOCK-ALIGNED tiles ---------------------------------------------------------------------------------
// The flat layout above puts candidates 2r, 2r+1 of the launch into tile row r, so a row's two halves can belong to
// different blocks and every half reads its own original (SAD) and its own H O H^T table row (SATD): 48 LDS.128 per row
// and tile. Here a block's k candidates occupy RPB = ceil(k / 2) consecutive tile rows (BPT = 120 / RPB blocks per
// tile), brought in by ONE 3-D TMA box per block: {128 bytes, RPB rows, 1 block} of the candidate array viewed as
// [blocks][RPB][128] with a block stride of 64 k bytes. (A box with a 64-byte inner dimension does not pack two
// candidates into a 128-byte row: under SWIZZLE_128B every box row starts a new 128-byte line.) For an odd k the box's
// last half-row is the first candidate of the NEXT block - a dead slot, never scored; only the very last block of the
// array has nothing behind it and is fetched as RPB - 1 rows plus its last candidate alone. Both halves of a row then
// share the original and the table row: 28 LDS.128 per row and tile, and one table read serves both halves' TMEM
// chunks. The price: that dead slot (4 % of the tile for k = 23). Measured on the bench's 8x8 level (k = 23, one 4K
// picture): 0.0624 ms flat -> 0.0550 ms.
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
               :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(s32addr(bar)) : "memory");
}

struct BaGeom { uint32_t n_blocks, k, rpb, bpt, recip; };   // recip = ceil(65536 / rpb): r / rpb = (r * recip) >> 16 for r < 128

__global__ void __launch_bounds__(kThreads, kCtasPerSm)
rough_tma_ba_kernel(const __grid_constant__ CUtensorMap map_cand, const __grid_constant__ CUtensorMap map_cand_m1,
                    const __grid_constant__ CUtensorMap map_c64, const __grid_constant__ CUtensorMap map_orig,
                    uint32_t *__restrict__ out, BaGeom gm)
{
  constexpr int OR = 8, kCandRows = kRows - OR;
  extern __shared__ __align__(1024) uint8_t dyn[];
  uint8_t *const sB = dyn;
  uint8_t *const sA = dyn + kSmemB;
  int *const sTO = reinterpret_cast<int *>(dyn + kSmemB + kStages * kTileBytes);     // [group][16][kTOPitch]
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], acc_full[2], acc_empty[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // B[n][kk], n = 64 s + c: the 64 pixels of half s of a row, sign (-1)^popcount(p & c), zero for the other half
  for (int i = tid; i < 128 * 8; i += kThreads) {
    const int n = i >> 3, ch = i & 7, s = n >> 6, c = n & 63;
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t v = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = 16 * ch + 4 * q + j, half = kk >> 6, p = kk & 63;
        const uint32_t b = half != s ? 0u : ((__popc(p & c) & 1) ? 0xffu : 0x01u);
        v |= b << (8 * j);
      }
      w[q] = v;
    }
    *reinterpret_cast<uint4 *>(sB + sw128(n, ch)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (tid == 0) {
    if (s32addr(dyn) & 1023) __trap();
    for (int i = 0; i < kStages; ++i) { bar_init(&full_bar[i], 1); bar_init(&empty_bar[i], 128); }
    for (int i = 0; i < 2; ++i) { bar_init(&acc_full[i], 1); bar_init(&acc_empty[i], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncwarp();
  if (warp == 9) tmem_alloc(&s_tmem, kTmemCols);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = s_tmem;
  const uint32_t n_tiles = (gm.n_blocks + gm.bpt - 1) / gm.bpt;
  const uint32_t tile0 = blockIdx.x, tstep = gridDim.x;
  const uint32_t my_tiles = tile0 < n_tiles ? (n_tiles - tile0 + tstep - 1) / tstep : 0;

  if (warp < 8) {
    // ===== epilogue: thread = tile row; two groups of 128 threads take alternate tiles =====
    const int r = tid & 127, grp = tid >> 7, wq = warp & 3;
    const uint32_t tlane = tmem + ((uint32_t)(wq * 32) << 16) + grp * kAccCols;
    const uint32_t blk = ((uint32_t)r * gm.recip) >> 16, m0 = 2 * ((uint32_t)r - blk * gm.rpb);
    const bool row_used = r < kCandRows && blk < gm.bpt;
    for (uint32_t i = grp; i < my_tiles; i += 2) {
      const uint32_t tile = tile0 + i * tstep, st = i % kStages;
      const uint8_t *tileA = sA + st * kTileBytes;
      int *const myTO = sTO + grp * (kMaxOrigRows * kTOPitch);
      const uint32_t b_first = tile * gm.bpt, gblk = b_first + blk;
      const uint32_t oidx = (b_first & 1) + blk;          // the tile's originals start at matrix row b_first >> 1
      const int torow = (int)(oidx >> 1), tohalf = (int)(oidx & 1);
      const bool blk_live = row_used && gblk < gm.n_blocks;
      const bool live[2] = { blk_live && m0 < gm.k, blk_live && m0 + 1 < gm.k };
      // ---- SAD first: it only needs the stage (the TMA's bytes), not the MMA -----------------------------------
      bar_wait(&full_bar[st], (i / kStages) & 1);
      unsigned sad[2];
      {
        const int orow = kCandRows + (row_used ? torow : 0);
        unsigned sd[2][2] = { { 0, 0 }, { 0, 0 } };
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 o = *reinterpret_cast<const uint4 *>(tileA + sw128(orow, 4 * tohalf + c));
          const uint4 p0 = *reinterpret_cast<const uint4 *>(tileA + sw128(r, c));
          const uint4 p1 = *reinterpret_cast<const uint4 *>(tileA + sw128(r, 4 + c));
          sd[0][0] = __vsadu4(p0.x, o.x) + sd[0][0]; sd[0][1] = __vsadu4(p0.y, o.y) + sd[0][1];
          sd[0][0] = __vsadu4(p0.z, o.z) + sd[0][0]; sd[0][1] = __vsadu4(p0.w, o.w) + sd[0][1];
          sd[1][0] = __vsadu4(p1.x, o.x) + sd[1][0]; sd[1][1] = __vsadu4(p1.y, o.y) + sd[1][1];
          sd[1][0] = __vsadu4(p1.z, o.z) + sd[1][0]; sd[1][1] = __vsadu4(p1.w, o.w) + sd[1][1];
        }
        sad[0] = sd[0][0] + sd[0][1]; sad[1] = sd[1][0] + sd[1][1];
      }
      bar_wait(&acc_full[i & 1], (i >> 1) & 1);
      fence_after_sync();
      bar_arrive(&empty_bar[st]);                              // stage read by this thread and consumed by the MMA: the TMA may refill it
      // ---- rows 120..127: H O H^T of the tile's blocks -> shared memory ------------------------------------------
      if (wq == 3) {
        int z[2][16];
        tmem_ld16(tlane, z[0]);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          tmem_ld_wait();
          if (q < 7) tmem_ld16(tlane + 16 * (q + 1), z[(q + 1) & 1]);
          if (r >= kCandRows) {
            int4 *dst = reinterpret_cast<int4 *>(myTO + (r - kCandRows) * kTOPitch + (q >> 2) * kTOHalf + 16 * (q & 3));
            const int(&v)[16] = z[q & 1];
            dst[0] = make_int4(v[0], v[1], v[2], v[3]); dst[1] = make_int4(v[4], v[5], v[6], v[7]);
            dst[2] = make_int4(v[8], v[9], v[10], v[11]); dst[3] = make_int4(v[12], v[13], v[14], v[15]);
          }
        }
      }
      // chunk order: coefficients 16 j .. 16 j + 15 of half 0, then of half 1 (TMEM columns 16 j and 64 + 16 j): one table
      // read serves both halves
      int z[2][16];
      tmem_ld16(tlane, z[0]);
      asm volatile("bar.sync %0, 128;\n" :: "r"(1 + grp) : "memory");
      const int *T = myTO + (row_used ? torow : 0) * kTOPitch + tohalf * kTOHalf;
      unsigned ac[2][4] = { { 0, 0, 0, 0 }, { 0, 0, 0, 0 } }, dc[2] = { 0, 0 };
      int4 t0, t1, t2, t3;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int j = q >> 1, h = q & 1;
        if (h == 0) {
          t0 = *reinterpret_cast<const int4 *>(T + 16 * j); t1 = *reinterpret_cast<const int4 *>(T + 16 * j + 4);
          t2 = *reinterpret_cast<const int4 *>(T + 16 * j + 8); t3 = *reinterpret_cast<const int4 *>(T + 16 * j + 12);
        }
        tmem_ld_wait();
        if (q < 7) {
          const int qn = q + 1;
          tmem_ld16(tlane + 64 * (qn & 1) + 16 * (qn >> 1), z[qn & 1]);
        } else {
          fence_before_sync();
          bar_arrive(&acc_empty[i & 1]);                      // accumulator drained: the next MMA into it may start
        }
        const int(&v)[16] = z[q & 1];
        unsigned(&a)[4] = ac[h];
        if (j == 0) dc[h] = (unsigned)abs(v[0] - t0.x);
        a[0] = __sad(v[0], t0.x, a[0]); a[1] = __sad(v[1], t0.y, a[1]); a[2] = __sad(v[2], t0.z, a[2]); a[3] = __sad(v[3], t0.w, a[3]);
        a[0] = __sad(v[4], t1.x, a[0]); a[1] = __sad(v[5], t1.y, a[1]); a[2] = __sad(v[6], t1.z, a[2]); a[3] = __sad(v[7], t1.w, a[3]);
        a[0] = __sad(v[8], t2.x, a[0]); a[1] = __sad(v[9], t2.y, a[1]); a[2] = __sad(v[10], t2.z, a[2]); a[3] = __sad(v[11], t2.w, a[3]);
        a[0] = __sad(v[12], t3.x, a[0]); a[1] = __sad(v[13], t3.y, a[1]); a[2] = __sad(v[14], t3.z, a[2]); a[3] = __sad(v[15], t3.w, a[3]);
      }
      const size_t oc = (size_t)gblk * gm.k + m0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const unsigned acc = (ac[h][0] + ac[h][1]) + (ac[h][2] + ac[h][3]);
        const unsigned satd = (acc - dc[h] + (dc[h] >> 2) + 2) >> 2;             // picture-generic.c:341-345
        if (live[h]) out[oc + h] = min(satd, 2 * sad[h]);
      }
    }
  } else if (warp == 8) {
    // ===== TMA producer (one lane): one 3-D box per block of the tile + the originals' rows =====
    if (lane == 0) {
      for (uint32_t i = 0; i < my_tiles; ++i) {
        const uint32_t st = i % kStages, tile = tile0 + i * tstep;
        if (i >= (uint32_t)kStages) bar_wait(&empty_bar[st], ((i / kStages) - 1) & 1);
        const uint32_t b_first = tile * gm.bpt, nb = min(gm.bpt, gm.n_blocks - b_first);
        const uint32_t dst = s32addr(sA + st * kTileBytes);
        // bytes: full boxes of RPB rows; the array's last block with an odd k: RPB - 1 rows + one 64-byte candidate
        const bool tail = (gm.k & 1) && b_first + nb == gm.n_blocks;
        bar_expect_tx(&full_bar[st], nb * gm.rpb * kRowBytes - (tail ? 64u : 0u) + OR * kRowBytes);
        for (uint32_t j = 0; j < nb; ++j) {
          const uint32_t d = dst + j * gm.rpb * kRowBytes;
          if (tail && j == nb - 1) {
            tma_load_3d(d, &map_cand_m1, 0, 0, (int)(b_first + j), &full_bar[st]);
            tma_load_2d(d + (gm.rpb - 1) * kRowBytes, &map_c64, 0, (int)((b_first + j) * gm.k + gm.k - 1), &full_bar[st]);
          } else {
            tma_load_3d(d, &map_cand, 0, 0, (int)(b_first + j), &full_bar[st]);
          }
        }
        tma_load_2d(dst + kCandRows * kRowBytes, &map_orig, 0, (int)(b_first >> 1), &full_bar[st]);
      }
    }
  } else if (lane == 0) {
    // ===== MMA issuer (warp 9, one lane) =====
    const uint64_t bd = smem_desc_sw128(s32addr(sB));
    for (uint32_t i = 0; i < my_tiles; ++i) {
      const uint32_t st = i % kStages;
      bar_wait(&full_bar[st], (i / kStages) & 1);
      if (i >= 2) bar_wait(&acc_empty[i & 1], ((i >> 1) - 1) & 1);
      fence_after_sync();
      const uint64_t ad = smem_desc_sw128(s32addr(sA + st * kTileBytes));
      const uint32_t d = tmem + (i & 1) * kAccCols;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_i8(d, ad + 2 * ks, bd + 2 * ks, kIdesc, ks > 0);   // 32 bytes of K per MMA
      umma_commit(&acc_full[i & 1]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 9) tmem_free(tmem, kTmemCols);
}


typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled()
{
  static std::once_flag once;
  static EncodeTiledFn fn = nullptr;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    cudaGetLastError();
  });
  return fn;
}

// rows x 128 bytes, box = 128 bytes x box_rows, SWIZZLE_128B; rows beyond the matrix read as zeros
static int make_map(CUtensorMap *m, const void *base, uint64_t rows, uint32_t box_rows)
{
  EncodeTiledFn f = encode_tiled();
  if (!f) return 1;
  const cuuint64_t dims[2] = { 128, rows }, strides[1] = { 128 };
  const cuuint32_t box[2] = { 128, box_rows }, estr[2] = { 1, 1 };
  const CUresult r = f(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}

// candidates viewed as [blocks][rows][128 bytes] with a block stride of 64 k bytes; box = `box_rows` rows of one block
static int make_map_cand3(CUtensorMap *m, const void *base, uint64_t n_blocks, uint32_t k, uint32_t rows, uint32_t box_rows)
{
  EncodeTiledFn f = encode_tiled();
  if (!f) return 1;
  const cuuint64_t dims[3] = { 128, rows, n_blocks }, strides[2] = { 128, 64ull * k };
  const cuuint32_t box[3] = { 128, box_rows, 1 }, estr[3] = { 1, 1, 1 };
  const CUresult r = f(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}

// single candidates: rows of 64 bytes, box = one row (lands at the start of a 128-byte line)
static int make_map_c64(CUtensorMap *m, const void *base, uint64_t n_cands)
{
  EncodeTiledFn f = encode_tiled();
  if (!f) return 1;
  const cuuint64_t dims[2] = { 64, n_cands }, strides[1] = { 64 };
  const cuuint32_t box[2] = { 64, 1 }, estr[2] = { 1, 1 };
  const CUresult r = f(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}

std::atomic<uint64_t> g_rough_tma_launches{0};

// Returns -1 when the layout does not qualify (caller falls through to the other kernels).
int launch_rough_tma(const Blocks &A, const Blocks &B, uint32_t *out, size_t n_items, int k, int n, cudaStream_t st)
{
  if (n != 8 && n != 16) return -1;
  if (k < 10) return -1;                                   // 16 original rows per tile must cover the tile's blocks
  const int orows = k >= 20 ? 8 : 16, crows = 128 - orows;
  if (B.stride != (uint32_t)n || B.bpr != 1 || B.row_pitch != (uint64_t)(n * n)) return -1;   // dense candidates
  if (A.stride != (uint32_t)n || A.bpr != 1 || A.row_pitch != (uint64_t)(n * n)) return -1;   // dense originals
  if ((reinterpret_cast<uintptr_t>(B.base) | reinterpret_cast<uintptr_t>(A.base)) & 127) return -1;
  const size_t n_blocks = n_items / (size_t)k;
  if (n_blocks * (size_t)k != n_items || n_items >= (1ull << 31)) return -1;
  if (n == 8 && (n_blocks & 1)) return -1;                 // rows hold pairs of 8x8 blocks
  CUtensorMap mc, mo;
  const uint64_t cand_rows = n_items * (uint64_t)(n * n) / 128, orig_rows = n_blocks * (uint64_t)(n * n) / 128;
  // 8x8 with 15..24 candidates per block: block-aligned tiles (UVG266_B200_ROUGH_TILES=flat keeps the flat layout)
  static const bool flat_only = getenv("UVG266_B200_ROUGH_TILES") && !strcmp(getenv("UVG266_B200_ROUGH_TILES"), "flat");
  if (n == 8 && k >= 15 && k <= 24 && !flat_only) {
    BaGeom gm;
    gm.n_blocks = (uint32_t)n_blocks; gm.k = (uint32_t)k; gm.rpb = (uint32_t)(k + 1) / 2; gm.bpt = 120 / gm.rpb;
    gm.recip = (65536 + gm.rpb - 1) / gm.rpb;
    CUtensorMap mc1, m64;
    if (make_map_cand3(&mc, B.base, n_blocks, (uint32_t)k, gm.rpb, gm.rpb) || make_map_cand3(&mc1, B.base, n_blocks, (uint32_t)k, gm.rpb, gm.rpb - 1) ||
        make_map_c64(&m64, B.base, n_items) || make_map(&mo, A.base, orig_rows, 8)) return -1;
    size_t tiles = (n_blocks + gm.bpt - 1) / gm.bpt;
    const size_t cap2 = (size_t)sm_count() * kCtasPerSm;
    if (tiles > cap2) tiles = cap2;
    static std::once_flag once_ba;
    static cudaError_t err_ba = cudaSuccess;
    std::call_once(once_ba, [] {
      err_ba = cudaFuncSetAttribute(reinterpret_cast<const void *>(rough_tma_ba_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
      if (err_ba == cudaSuccess)
        err_ba = cudaFuncSetAttribute(reinterpret_cast<const void *>(rough_tma_ba_kernel), cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared);
    });
    UVGB_TRY(err_ba);
    rough_tma_ba_kernel<<<(unsigned)tiles, kThreads, kSmemTotal, st>>>(mc, mc1, m64, mo, out, gm);
    UVGB_LAUNCH_CHECK();
    g_rough_tma_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
  }
  if (make_map(&mc, B.base, cand_rows, (uint32_t)crows) || make_map(&mo, A.base, orig_rows, (uint32_t)orows)) return -1;
  FastDivK kd;
  kd.k = (uint32_t)k;
  kd.magic = k > 1 ? (~0ull / (uint64_t)k) + 1 : 0;
  const size_t per_tile = n == 8 ? 2 * crows : crows / 2;
  size_t grid = (n_items + per_tile - 1) / per_tile;
  const size_t cap = (size_t)sm_count() * kCtasPerSm;
  if (grid > cap) grid = cap;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    // two CTAs (2 x ~99 KB) per SM need the full shared-memory carve-out
    auto setup = [](const void *fn) {
      cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      return e;
    };
    attr_err = setup(reinterpret_cast<const void *>(rough_tma_kernel<8, 8>));
    if (attr_err == cudaSuccess) attr_err = setup(reinterpret_cast<const void *>(rough_tma_kernel<8, 16>));
    if (attr_err == cudaSuccess) attr_err = setup(reinterpret_cast<const void *>(rough_tma_kernel<16, 8>));
    if (attr_err == cudaSuccess) attr_err = setup(reinterpret_cast<const void *>(rough_tma_kernel<16, 16>));
  });
  UVGB_TRY(attr_err);
  if (n == 8) {
    if (orows == 8) rough_tma_kernel<8, 8><<<(unsigned)grid, kThreads, kSmemTotal, st>>>(mc, mo, out, (uint32_t)n_items, kd);
    else rough_tma_kernel<8, 16><<<(unsigned)grid, kThreads, kSmemTotal, st>>>(mc, mo, out, (uint32_t)n_items, kd);
  } else {
    if (orows == 8) rough_tma_kernel<16, 8><<<(unsigned)grid, kThreads, kSmemTotal, st>>>(mc, mo, out, (uint32_t)n_items, kd);
    else rough_tma_kernel<16, 16><<<(unsigned)grid, kThreads, kSmemTotal, st>>>(mc, mo, out, (uint32_t)n_items, kd);
  }
  UVGB_LAUNCH_CHECK();
  g_rough_tma_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // namespace uvgb

// launches of the tcgen05 + TMA rough-cost kernel since load (tests assert that it, not a fallback, ran)
extern "C" uint64_t uvg_b200_rough_tma_launches(void) { return uvgb::g_rough_tma_launches.load(); }
