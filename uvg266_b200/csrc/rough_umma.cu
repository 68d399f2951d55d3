// rough_umma.cu - rough-search cost min(SATD, 2*SAD), Hadamard on the 5th-gen tensor cores
// (tcgen05.mma kind::i8, accumulators in TMEM).
//
// Semantics: search_intra.c:149-163 (get_cost) over the satd/sad NxN duals
// (picture-generic.c:256-348 8x8 Hadamard incl. the |dc|>>2 rule, :371-410, :1091-1113).
//
// The packed-ALU Hadamard (rough.cu) needs ~860 issue slots per 8x8 candidate and saturates the
// issue ports at 40-43 % of the HBM roofline; mma.sync IMMA (rough_mma.cu) keeps the accumulators
// in registers and pays ~390 instructions of epilogue per tile, ending up no faster. tcgen05 does
// the same product with the accumulators in TMEM:
//   the 2-D Hadamard of an 8x8 block is one product with the 64x64 +-1 matrix H8 (x) H8
//   (entry (k, n) = (-1)^popcount(k & n)); H(O - P)H^T = H O H^T - H P H^T, so the u8 pixels are
//   the A operand as they are, and the original block is just one more row of the same tile.
// One CTA tile = 128 rows x K = 128 bytes: row = one 8x8 sub-block of one candidate, bytes 0..63
// its 64 predicted pixels, bytes 64..127 the 64 original pixels of the same sub-block, and
// B = [H8 (x) H8 ; -(H8 (x) H8)] (64 x 128 s8), so D = P H - O H = (P - O) H comes out of the tensor
// core directly: every row is self-contained (any k, no tile/block alignment) and there is no
// exchange between rows.
// Warp-specialised CTA (9 warps), no __syncthreads in the steady state, mbarriers only:
//   warps 0-7        also the producers: every thread cp.asyncs (LDGSTS) 4 candidate chunks and its
//                    row's 4 original chunks of the group's NEXT tile straight into the canonical
//                    K-major SWIZZLE_128B layout (chunk index XOR row & 7) of a 4-stage ring (the
//                    originals are re-read through L1/L2 for each candidate, HBM sees them once);
//                    completion: cp.async.mbarrier.arrive.noinc - no wait_group, no fence;
//   warp 8 (1 lane)  four tcgen05.mma.cta_group::1.kind::i8 (M128 N64 K32) per tile into one of
//                    two 64-column TMEM accumulators, tcgen05.commit -> mbarrier;
//   warps 0-7        two epilogue groups taking alternate tiles, one thread per row: SAD from its own swizzled row (conflict-free
//                    LDS.128), tcgen05.ld 64 coefficients, sum |c| with the |dc| >> 2 rule,
//                    shuffle-reduce over the sub-blocks of a candidate, one store.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>
#include <stdio.h>

namespace uvgb {

namespace {

__device__ __forceinline__ uint32_t s32addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bar_init(uint64_t *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(s32addr(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(s32addr(bar)) : "memory");
}
// Bounded wait: a descriptor mistake must end in a trap, not in a hung GPU.
__device__ __forceinline__ void bar_wait(uint64_t *bar, uint32_t parity)
{
  uint32_t ok = 0;
  // the suspend-time hint parks the warp in the barrier unit instead of spinning through issue slots
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

// D[tmem] (+)= A[smem desc] * B[smem desc], u8 x s8 -> s32, issued by ONE thread.
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
// 16 consecutive TMEM columns of this thread's lane (warp w reads lanes 32w..32w+31)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16])
{
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start address >> 4 [0,14), LBO >> 4 [16,30) (unused: K fits one swizzle atom), SBO >> 4 [32,46)
// = 8 rows x 128 B, version 1 [46,48), layout type SWIZZLE_128B = 2 [61,64).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr)
{
  return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (2 << 4), A = u8 (0 << 7), B = s8
// (1 << 10), both K-major, N = 64 (N >> 3 << 17), M = 128 (M >> 4 << 24).
constexpr uint32_t kIdesc = (2u << 4) | (0u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

constexpr int kTileRows = 128, kRowBytes = 128, kTileBytes = kTileRows * kRowBytes, kTmemCols = 128, kCtasPerSm = 3,
              kStages = 4,    // tiles in the shared-memory ring; two TMEM accumulators of 64 columns
              kThreads = 288;    // 2 epilogue groups of 4 warps (they also issue the copies) + MMA warp

__device__ __forceinline__ int sw128(int row, int chunk) { return row * kRowBytes + ((chunk ^ (row & 7)) << 4); }

struct FastDivK { uint32_t k; uint64_t magic; };      // floor(x / k) for x < 2^32: (x * ceil(2^64 / k)) >> 64
__device__ __forceinline__ uint32_t fdivk(uint32_t x, const FastDivK &d)
{
  return d.k == 1 ? x : (uint32_t)__umul64hi((uint64_t)x, d.magic);
}

}  // namespace


template <int N, bool DENSE>      // DENSE: the original blocks are dense N x N, one after the other
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
rough_umma_kernel(Blocks A, const uint8_t *__restrict__ cand, uint32_t *__restrict__ out, uint32_t n_units, FastDivK kd)
{
  constexpr int NS = N / 8, NSUB = NS * NS, LOG_NSUB = N == 8 ? 0 : N == 16 ? 2 : 4;
  extern __shared__ __align__(1024) uint8_t dyn_smem[];
  uint8_t *const sB = dyn_smem;                                  // 64 x 128 s8
  uint8_t *const sA = dyn_smem + 64 * kRowBytes;                 // kStages tiles of 128 x 128 u8
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], acc_full[2], acc_empty[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // B[n][kk] = H64[kk][n] for kk < 64 (predicted pixels), -H64[kk - 64][n] for the original's pixels
  for (int i = tid; i < 64 * 8; i += kThreads) {
    const int n = i >> 3, c = i & 7;
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t v = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = 16 * c + 4 * q + j;
        const int neg = (__popc((kk & 63) & n) + (kk >> 6)) & 1;
        v |= (neg ? 0xffu : 0x01u) << (8 * j);
      }
      w[q] = v;
    }
    *reinterpret_cast<uint4 *>(sB + sw128(n, c)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) { bar_init(&full_bar[i], 128); bar_init(&empty_bar[i], 128); }
    for (int i = 0; i < 2; ++i) { bar_init(&acc_full[i], 1); bar_init(&acc_empty[i], 128); }
  }
  __syncwarp();
  if (warp == 8) tmem_alloc(&s_tmem, kTmemCols);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = s_tmem;
  const uint32_t n_tiles = (n_units + kTileRows - 1) / kTileRows;
  const uint32_t tile0 = blockIdx.x, tstep = gridDim.x;
  const uint32_t my_tiles = tile0 < n_tiles ? (n_tiles - tile0 + tstep - 1) / tstep : 0;

  if (warp < 8) {
    // ===== epilogue: thread = tile row =====
    // Each thread also copies the original's 64 bytes of ITS row for the tile kAhead iterations ahead
    // (4 cp.async; neighbouring rows mostly share the block, so a warp's copies coalesce in L1).
    // Two groups of 128 threads take alternate tiles (group g: tiles g, g+2, ...; accumulator g), so
    // one group's TMEM read / sums overlap the other's waits: the epilogue is a long dependent
    // instruction stream and needs the extra warps per scheduler.
    const int r = tid & 127, grp = tid >> 7;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16) + grp * 64;
    const uint32_t o_sub = r & (NSUB - 1), o_sy = o_sub / NS, o_sx = o_sub - o_sy * NS;
    const uint32_t o_stride = DENSE ? N : A.stride;
    const size_t o_pix = (size_t)(o_sy * 8) * o_stride + o_sx * 8;
    // per-thread constants of the copies: own row's original chunks 4..7, and candidate chunk r & 3 of
    // the four rows (r & ~3) | c (lanes 4m..4m+3 copy 64 contiguous bytes of one row)
    const uint32_t kc = r & 3;
    uint32_t od[4], cd[4], coff[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      od[c] = r * kRowBytes + (((4 + c) ^ (r & 7)) << 4);
      const uint32_t row = (r & ~3u) | c, sub = row & (NSUB - 1), sy = sub / NS, sx = sub - sy * NS;
      cd[c] = row * kRowBytes + ((kc ^ (row & 7)) << 4);
      coff[c] = (row >> LOG_NSUB) * (N * N) + (sy * 8 + 2 * kc) * N + sx * 8;     // inside the tile's 8 KB of candidates
    }
    const uint32_t sA_addr = s32addr(sA);
    auto fetch_orig = [&](uint32_t j) {                 // j = this CTA's tile counter
      const uint32_t st = j % kStages;
      if (j < my_tiles) {
        if (j >= (uint32_t)kStages) bar_wait(&empty_bar[st], ((j / kStages) - 1) & 1);
        const uint32_t tile = tile0 + j * tstep, u = tile * kTileRows + r, dbase = sA_addr + st * kTileBytes;
        const uint8_t *ctile = cand + (size_t)tile * (kTileRows * 64);            // 128 sub-blocks x 64 bytes
        const bool full = (tile + 1) * kTileRows <= n_units;                      // every tile but the last
        if (full || u < n_units) {
          const uint32_t b = fdivk(u >> LOG_NSUB, kd);
          const uint8_t *p = static_cast<const uint8_t *>(A.base) + (DENSE ? (size_t)b * (N * N) : (size_t)A.offset(b)) + o_pix;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (N == 8 && DENSE) cp_async16(dbase + od[c], p + 16 * c);           // two pixel rows per 16-byte copy
            else { cp_async8(dbase + od[c], p + (size_t)(2 * c) * o_stride); cp_async8(dbase + od[c] + 8, p + (size_t)(2 * c + 1) * o_stride); }
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (full || tile * kTileRows + ((r & ~3u) | c) < n_units) {
            if (N == 8) cp_async16(dbase + cd[c], ctile + coff[c]);
            else { cp_async8(dbase + cd[c], ctile + coff[c]); cp_async8(dbase + cd[c] + 8, ctile + coff[c] + N); }
          }
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" :: "r"(s32addr(&full_bar[st])) : "memory");
      }
    };
    fetch_orig(grp);
    for (uint32_t i = grp; i < my_tiles; i += 2) {
      const uint32_t tile = tile0 + i * tstep, st = i % kStages, u = tile * kTileRows + r;
      const bool live = u < n_units;
      const uint8_t *tileA = sA + st * kTileBytes;
      fetch_orig(i + 2);                                 // this group's next tile
      bar_wait(&full_bar[st], (i / kStages) & 1);
      unsigned sd[4] = { 0, 0, 0, 0 };
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 p = *reinterpret_cast<const uint4 *>(tileA + (od[c] ^ 64));       // chunk c = chunk 4 + c with bit 2 of the index flipped
        const uint4 o = *reinterpret_cast<const uint4 *>(tileA + od[c]);
        sd[0] = __vsadu4(p.x, o.x) + sd[0]; sd[1] = __vsadu4(p.y, o.y) + sd[1];
        sd[2] = __vsadu4(p.z, o.z) + sd[2]; sd[3] = __vsadu4(p.w, o.w) + sd[3];
      }
      unsigned sad = (sd[0] + sd[1]) + (sd[2] + sd[3]);
      bar_wait(&acc_full[i & 1], (i >> 1) & 1);
      fence_after_sync();
      unsigned ac[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }, dc = 0;
#pragma unroll
      for (int half = 0; half < 2; ++half) {              // 2 x 32 columns: keeps the thread under 68 registers
        int z[2][16];
        tmem_ld16(tlane + 32 * half, z[0]);
        tmem_ld16(tlane + 32 * half + 16, z[1]);
        tmem_ld_wait();
        if (half == 1) {
          fence_before_sync();
          bar_arrive(&acc_empty[i & 1]);                  // accumulator drained
          bar_arrive(&empty_bar[st]);                     // stage read (SAD) and consumed by the MMA
        } else {
          dc = abs(z[0][0]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int j = 0; j < 16; ++j) ac[j & 7] = __sad(z[h][j], 0, ac[j & 7]);
      }
      const unsigned acc = ((ac[0] + ac[1]) + (ac[2] + ac[3])) + ((ac[4] + ac[5]) + (ac[6] + ac[7]));
      unsigned satd = live ? (acc - dc + (dc >> 2) + 2) >> 2 : 0;     // picture-generic.c:341-345
      if (!live) sad = 0;
      if (NSUB > 1) { satd = group_sum<NSUB>(satd); sad = group_sum<NSUB>(sad); }
      if (live && (u & (NSUB - 1)) == 0) out[u >> LOG_NSUB] = min(satd, 2 * sad);
    }
  } else if (warp == 8 && lane == 0) {
    // ===== MMA issuer (one thread) =====
    const uint64_t bd = smem_desc_sw128(s32addr(sB));
    for (uint32_t i = 0; i < my_tiles; ++i) {
      const uint32_t st = i % kStages;
      bar_wait(&full_bar[st], (i / kStages) & 1);
      if (i >= 2) bar_wait(&acc_empty[i & 1], ((i >> 1) - 1) & 1);
      fence_async_smem();                               // cp.async wrote the stage through the generic proxy
      fence_after_sync();
      const uint64_t ad = smem_desc_sw128(s32addr(sA + st * kTileBytes));
      const uint32_t d = tmem + (i & 1) * 64;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_i8(d, ad + 2 * ks, bd + 2 * ks, kIdesc, ks > 0);   // 32 bytes of K per MMA
      umma_commit(&acc_full[i & 1]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_free(tmem, kTmemCols);
}

// Returns -1 when the layout does not qualify (caller falls through to the other kernels).
int launch_rough_umma(const Blocks &A, const Blocks &B, uint32_t *out, size_t n_items, int k, int n, cudaStream_t st)
{
  if (n != 8 && n != 16 && n != 32) return -1;
  if (B.stride != (uint32_t)n || B.bpr != 1 || B.row_pitch != (uint64_t)(n * n)) return -1;   // dense candidates
  if (reinterpret_cast<uintptr_t>(B.base) & 15) return -1;
  if ((reinterpret_cast<uintptr_t>(A.base) & 7) || (A.stride & 7) || (A.col_pitch & 7) || (A.row_pitch & 7)) return -1;
  const size_t n_blocks = n_items / (size_t)k;
  if (n_blocks * (size_t)k != n_items || n_blocks >= (1ull << 31)) return -1;
  const int nsub = (n / 8) * (n / 8);
  if (n_items * (size_t)nsub >= (1ull << 32)) return -1;
  const uint32_t n_units = (uint32_t)(n_items * (size_t)nsub);
  FastDivK kd;
  kd.k = (uint32_t)k;
  kd.magic = k > 1 ? (~0ull / (uint64_t)k) + 1 : 0;
  // persistent CTAs: exactly one resident wave; __launch_bounds__(128, kCtasPerSm), 40 KB of shared
  // memory and 64 TMEM columns per CTA
  size_t grid = (n_units + kTileRows - 1) / kTileRows;
  const size_t cap = (size_t)sm_count() * kCtasPerSm;
  if (grid > cap) grid = cap;
  const uint8_t *cb = static_cast<const uint8_t *>(B.base);
  constexpr int smem = 64 * kRowBytes + kStages * kTileBytes;      // 56 KB: opt in once per kernel
  const bool dense = A.bpr == 1 && A.stride == (uint32_t)n && A.row_pitch == (uint64_t)(n * n);
  static bool attr_set = false;
  if (!attr_set) {
#define UVGB_UMMA_ATTR(NN, DD) \
    UVGB_TRY(cudaFuncSetAttribute(rough_umma_kernel<NN, DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))
    UVGB_UMMA_ATTR(8, true); UVGB_UMMA_ATTR(8, false); UVGB_UMMA_ATTR(16, true); UVGB_UMMA_ATTR(16, false);
    UVGB_UMMA_ATTR(32, true); UVGB_UMMA_ATTR(32, false);
#undef UVGB_UMMA_ATTR
    attr_set = true;
  }
#define UVGB_UMMA_GO(NN, DD) rough_umma_kernel<NN, DD><<<(unsigned)grid, kThreads, smem, st>>>(A, cb, out, n_units, kd)
  if (n == 8) { if (dense) UVGB_UMMA_GO(8, true); else UVGB_UMMA_GO(8, false); }
  else if (n == 16) { if (dense) UVGB_UMMA_GO(16, true); else UVGB_UMMA_GO(16, false); }
  else { if (dense) UVGB_UMMA_GO(32, true); else UVGB_UMMA_GO(32, false); }
#undef UVGB_UMMA_GO
  UVGB_LAUNCH_CHECK();
  return 0;
}

}  // namespace uvgb
