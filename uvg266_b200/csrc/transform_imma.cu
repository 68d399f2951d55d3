// transform_imma.cu - the standalone 2-D transforms of square 16x16 / 32x32 TUs on the INT8 tensor cores
// (mma.sync.m16n8k16, s32 accumulate): what uvg_b200_transform_dev / the dct_NxN / mts_dct drop-ins run when a
// configuration does not go through the fused quantize_residual kernels. Reference: dct-generic.c:2560-2678
// (mts_dct_generic / mts_idct_generic and the fastForward* / fastInverse* kernels they dispatch to); bit-exact.
//
// One warp transforms one TU entirely in registers. A 16-bit value is multiplied as two byte planes (lo u8, hi s8)
// with separate int32 accumulators recombined as accH*256 + accL - exact: every |sum| <= 32 * 32768 * 90 < 2^27.
// The C fragment of the first product is the B (forward) or A (inverse) operand of the second one through the
// contraction-slot permutation pi of imma.cuh, so nothing is staged between the passes:
//   forward   P1  D' = Th * X^T        A = Th  (const),  B = rows of X as loaded (4 consecutive int16 per lane)
//             P2  Y  = Tv * D'^T       A = Tv  (const, contraction pi),  B = D' fragments, byte split
//                 (the rows of Th in P1 are permuted - sigma_row - so that a lane ends up with 4 consecutive
//                 frequencies of a row of Y: 8-byte stores)
//   inverse   P1  U  = Tv^T * C        A = Tv^T (const, contraction pi), B = ldmatrix.trans of the coefficient block
//                                      staged once in shared memory (the contraction index kv runs down its columns)
//             P2  R  = U * Th          A = U fragments, byte split,  B = Th (const, contraction pi, output columns
//                                      permuted so that a lane ends up with 4 consecutive samples: 8-byte stores)
// Rounding shifts, the int16 wrap (forward) / clip (inverse) and the zero-out of dct-generic.c run on the int32
// accumulators between the products.
#include "transform.cuh"
#include "imma.cuh"
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace uvgb {
namespace {

// Row (g, r) of a 16-row A tile of the forward P1 carries frequency sigma_row: the C fragment rows of P1 are the B
// columns of P2, whose C fragment then holds, over a pair of 8-wide tiles, FOUR consecutive frequencies per lane
// (8-byte stores, 32 contiguous bytes per row and quad: whole sectors).
__host__ __device__ constexpr int sigma_row(int g, int r) { return 4 * (g >> 1) + 2 * r + (g & 1); }

template <int N> struct TrShape {
  static constexpr int MT = N / 16, NT = N / 8, KT = N / 16;
  static constexpr int NA = MT * KT * 2 * 32, NB = KT * NT * 32;
};

// Constant operand fragments in the register / lane order mma.sync expects:
//   [0,NA)        AH   P1 fwd  A = Th      natural contraction order (sample columns), rows permuted by sigma_row
//   [NA,2NA)      AV   P2 fwd  A = Tv      contraction permuted by pi
//   [2NA,3NA)     AVT  P1 inv  A = Tv^T    contraction permuted by pi
//   [3NA,3NA+NB)  BH   P2 inv  B = Th      contraction pi, output columns sigma
template <int N>
void build(const int8_t *th, const int8_t *tv, std::vector<uint32_t> &out)
{
  using S = TrShape<N>;
  out.assign(3 * S::NA + S::NB, 0);
  for (int e0 = 0; e0 < S::NA; ++e0) {
    const int ln = e0 & 31, r = (e0 >> 5) & 1, tile = e0 >> 6, kt = tile % S::KT, mt = tile / S::KT;
    const int g = ln >> 2, t = ln & 3, row = mt * 16 + g + 8 * r;
    int ah[4], av[4], avt[4];
    for (int e = 0; e < 4; ++e) {
      ah[e] = th[(mt * 16 + sigma_row(g, r)) * N + kt * 16 + 4 * t + e];
      av[e] = tv[row * N + kt * 16 + pi_col(t, e)];
      avt[e] = tv[(kt * 16 + pi_col(t, e)) * N + row];
    }
    out[e0] = pack4(ah[0], ah[1], ah[2], ah[3]);
    out[S::NA + e0] = pack4(av[0], av[1], av[2], av[3]);
    out[2 * S::NA + e0] = pack4(avt[0], avt[1], avt[2], avt[3]);
  }
  for (int e0 = 0; e0 < S::NB; ++e0) {
    const int ln = e0 & 31, tile = e0 >> 5, nt = tile % S::NT, kt = tile / S::NT;
    const int g = ln >> 2, t = ln & 3;
    const int ih = 16 * (nt >> 1) + 4 * (g >> 1) + 2 * (nt & 1) + (g & 1);      // sigma
    int bh[4];
    for (int e = 0; e < 4; ++e) bh[e] = th[(kt * 16 + pi_col(t, e)) * N + ih];
    out[3 * S::NA + e0] = pack4(bh[0], bh[1], bh[2], bh[3]);
  }
}

const uint32_t *tables(int N, int th, int tv)
{
  static std::mutex mu;
  static const uint32_t *cache[2][3][3] = {};
  const int si = N == 16 ? 0 : 1;
  std::lock_guard<std::mutex> lk(mu);
  if (cache[si][th][tv]) return cache[si][th][tv];
  std::vector<uint32_t> h;
  if (N == 16) build<16>(tr_matrix_host(th, N), tr_matrix_host(tv, N), h); else build<32>(tr_matrix_host(th, N), tr_matrix_host(tv, N), h);
  uint32_t *d = nullptr;
  if (cudaMalloc(&d, h.size() * 4) != cudaSuccess) return nullptr;
  if (cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  cache[si][th][tv] = d;
  return d;
}

struct ImmaTr {
  const uint32_t *tables;
  int shift1, shift2;
  int cut_a, cut_b, cut_c;   // forward: kh >= cut_a zero after P1; kv >= cut_b or kh >= cut_c zero after P2
                             // inverse: kh >= cut_a zero after P1
};

template <int N>
__global__ void __launch_bounds__(128, N == 32 ? 4 : 10)
transform_imma_fwd_kernel(const int16_t *__restrict__ in, int16_t *__restrict__ out, size_t n_tus, ImmaTr p)
{
  using S = TrShape<N>;
  constexpr int MT = S::MT, NT = S::NT, KT = S::KT;
  __shared__ uint32_t sAH[S::NA], sAV[S::NA];
  for (int e = threadIdx.x; e < S::NA; e += blockDim.x) { sAH[e] = p.tables[e]; sAV[e] = p.tables[S::NA + e]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5);
  const int add1 = p.shift1 > 0 ? 1 << (p.shift1 - 1) : 0, add2 = 1 << (p.shift2 - 1);
  // rows j = nt*8 + g, samples kt*16 + 4t .. +3 of a TU; the next TU's words are requested before this one's products
  // start (one warp per TU and a dependent load -> compute -> store chain would otherwise leave too few bytes in flight)
  uint2 w[NT][KT];
  auto fetch = [&](size_t tu) {
    const int16_t *src = in + tu * (size_t)(N * N);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) w[nt][kt] = __ldcs(reinterpret_cast<const uint2 *>(src + (nt * 8 + g) * N + kt * 16 + 4 * t));
  };
  const size_t tu0 = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tu0 < n_tus) fetch(tu0);
  for (size_t tu = tu0; tu < n_tus; tu += warps) {
    int16_t *dst = out + tu * (size_t)(N * N);
    uint32_t lo[NT][KT], hi[NT][KT];                          // byte planes
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        lo[nt][kt] = __byte_perm(w[nt][kt].x, w[nt][kt].y, 0x6420);
        hi[nt][kt] = __byte_perm(w[nt][kt].x, w[nt][kt].y, 0x7531);
      }
    if (tu + warps < n_tus) fetch(tu + warps);
    // ---- P1: D'[kh][j] --------------------------------------------------------------------------------------------
    int v[MT][NT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        int aL[4] = { 0, 0, 0, 0 }, aH[4] = { 0, 0, 0, 0 };
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
          const int fi = ((mt * KT + kt) * 2) * 32 + lane;
          imma<false, true>(aL, sAH[fi], sAH[fi + 32], lo[nt][kt]);
          imma<false, false>(aH, sAH[fi], sAH[fi + 32], hi[nt][kt]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int kh = mt * 16 + sigma_row(g, c >> 1);
          const int s = (int)(short)((aH[c] * 256 + aL[c] + add1) >> p.shift1);
          v[mt][nt][c] = kh < p.cut_a ? s : 0;
        }
      }
    // ---- P2: Y[kv][kh] ----------------------------------------------------------------------------------------------
#pragma unroll
    for (int q = 0; q < NT / 2; ++q) {                        // 16 frequencies kh = 16q + 4t + e (two 8-wide B tiles)
      uint32_t bl[2][KT], bh[2][KT];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int kt = 0; kt < KT; ++kt)
          split4(v[q][2 * kt][2 * r], v[q][2 * kt][2 * r + 1], v[q][2 * kt + 1][2 * r], v[q][2 * kt + 1][2 * r + 1], bl[r][kt], bh[r][kt]);
#pragma unroll
      for (int mt2 = 0; mt2 < MT; ++mt2) {                    // 16 frequencies kv
        int aL[2][4], aH[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
          for (int c = 0; c < 4; ++c) { aL[r][c] = 0; aH[r][c] = 0; }
#pragma unroll
          for (int kt = 0; kt < KT; ++kt) {
            const int fi = ((mt2 * KT + kt) * 2) * 32 + lane;
            imma<false, true>(aL[r], sAV[fi], sAV[fi + 32], bl[r][kt]);
            imma<false, false>(aH[r], sAV[fi], sAV[fi + 32], bh[r][kt]);
          }
        }
        const int kh = q * 16 + 4 * t;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {                      // C rows g, g + 8
          const int kv = mt2 * 16 + g + 8 * rr;
          int y[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = e >> 1, c = 2 * rr + (e & 1);
            y[e] = (int)(short)((aH[r][c] * 256 + aL[r][c] + add2) >> p.shift2);
            if (kv >= p.cut_b || kh + e >= p.cut_c) y[e] = 0;
          }
          __stcs(reinterpret_cast<uint2 *>(dst + kv * N + kh),
                 make_uint2(__byte_perm((uint32_t)y[0], (uint32_t)y[1], 0x5410), __byte_perm((uint32_t)y[2], (uint32_t)y[3], 0x5410)));
        }
      }
    }
  }
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t smem_addr)
{
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_addr));
}

template <int N>
__global__ void __launch_bounds__(128, N == 32 ? 4 : 10)
transform_imma_inv_kernel(const int16_t *__restrict__ in, int16_t *__restrict__ out, size_t n_tus, ImmaTr p)
{
  using S = TrShape<N>;
  constexpr int MT = S::MT, NT = S::NT, KT = S::KT;
  constexpr int P = N * 2 + 16;                               // row pitch of the staged block, bytes: 8 rows of an
                                                              // ldmatrix tile fall into 8 different 16-byte bank groups
  __shared__ uint32_t sAVT[S::NA], sBH[S::NB];
  __shared__ __align__(16) uint8_t sC[4][N * P];
  for (int e = threadIdx.x; e < S::NA; e += blockDim.x) sAVT[e] = p.tables[2 * S::NA + e];
  for (int e = threadIdx.x; e < S::NB; e += blockDim.x) sBH[e] = p.tables[3 * S::NA + e];
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, warp = threadIdx.x >> 5;
  const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5);
  const int add1 = 1 << (p.shift1 - 1), add2 = 1 << (p.shift2 - 1);
  uint8_t *mine = sC[warp];
  const uint32_t mine_s = (uint32_t)__cvta_generic_to_shared(mine);
  constexpr int VPR = N / 8, VPL = N * VPR / 32;              // 16-byte vectors per row / per lane
  uint4 w[VPL];                                               // the next TU's coefficients, requested one TU ahead
  auto fetch = [&](size_t tu) {
    const uint4 *src = reinterpret_cast<const uint4 *>(in + tu * (size_t)(N * N));
#pragma unroll
    for (int i = 0; i < VPL; ++i) w[i] = __ldcs(src + lane + 32 * i);
  };
  const size_t tu0 = (size_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (tu0 < n_tus) fetch(tu0);
  for (size_t tu = tu0; tu < n_tus; tu += warps) {
    int16_t *dst = out + tu * (size_t)(N * N);
    __syncwarp();                                             // previous block's ldmatrix reads
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = lane + 32 * i, row = c / VPR, col = c % VPR;
      *reinterpret_cast<uint4 *>(mine + row * P + col * 16) = w[i];
    }
    __syncwarp();
    if (tu + warps < n_tus) fetch(tu + warps);
    // ---- P1: U[y][kh] = sum_kv Tv[kv][y] * C[kv][kh], shift 7, clip -----------------------------------------------------
    int u[MT][NT][4];
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {                     // a pair of 8-wide kh tiles per ldmatrix.x4
      int aL[2][MT][4], aH[2][MT][4];
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int c = 0; c < 4; ++c) { aL[q][mt][c] = 0; aH[q][mt][c] = 0; }
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        // matrices: 0 = (kv 16kt..+7, tile 2np), 1 = (kv 16kt+8.., tile 2np), 2 / 3 = the same rows of tile 2np+1
        uint32_t r[4];
        const int m = lane >> 3, rr = lane & 7;
        ldmatrix_x4_trans(r, mine_s + (16 * kt + 8 * (m & 1) + rr) * P + (8 * (2 * np + (m >> 1))) * 2);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          // slots: (kv 2t, 2t+1 of the upper tile, kv 8+2t, 8+2t+1 of the lower one) = pi
          const uint32_t bl = __byte_perm(r[2 * q], r[2 * q + 1], 0x6420), bh = __byte_perm(r[2 * q], r[2 * q + 1], 0x7531);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const int fi = ((mt * KT + kt) * 2) * 32 + lane;
            imma<false, true>(aL[q][mt], sAVT[fi], sAVT[fi + 32], bl);
            imma<false, false>(aH[q][mt], sAVT[fi], sAVT[fi + 32], bh);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int nt = 2 * np + q, kh = nt * 8 + 2 * t + (c & 1);
            const int s = clip3(-32768, 32767, (aH[q][mt][c] * 256 + aL[q][mt][c] + add1) >> p.shift1);
            u[mt][nt][c] = kh < p.cut_a ? s : 0;
          }
    }
    // ---- P2: R[y][x] = sum_kh U[y][kh] * Th[kh][x], shift 20 - bitdepth, clip -----------------------------------------
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      int aL[NT][4], aH[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) { aL[nt][c] = 0; aH[nt][c] = 0; }
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        uint32_t lo0, hi0, lo1, hi1;
        split4(u[mt][2 * kt][0], u[mt][2 * kt][1], u[mt][2 * kt + 1][0], u[mt][2 * kt + 1][1], lo0, hi0);   // row g
        split4(u[mt][2 * kt][2], u[mt][2 * kt][3], u[mt][2 * kt + 1][2], u[mt][2 * kt + 1][3], lo1, hi1);   // row g + 8
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const uint32_t b0 = sBH[(kt * NT + nt) * 32 + lane];
          imma<true, false>(aL[nt], lo0, lo1, b0);
          imma<false, false>(aH[nt], hi0, hi1, b0);
        }
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int qs = 0; qs < KT; ++qs) {
          int x[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int nt = 2 * qs + (e >> 1), c = 2 * hh + (e & 1);
            x[e] = clip3(-32768, 32767, (aH[nt][c] * 256 + aL[nt][c] + add2) >> p.shift2);
          }
          const uint2 w = make_uint2(__byte_perm((uint32_t)x[0], (uint32_t)x[1], 0x5410), __byte_perm((uint32_t)x[2], (uint32_t)x[3], 0x5410));
          __stcs(reinterpret_cast<uint2 *>(dst + (mt * 16 + g + 8 * hh) * N + qs * 16 + 4 * t), w);
        }
    }
  }
}

}  // namespace

// 0: launched; -1: shape / alignment not covered (the caller falls back); 1: error
int launch_transform_imma(bool inverse, const TrPlan &p, const int16_t *in, int16_t *out, size_t n_tus, cudaStream_t st)
{
  static const bool off = getenv("UVG266_B200_TRANSFORM") && !strcmp(getenv("UVG266_B200_TRANSFORM"), "fast");
  if (off || p.w != p.h || (p.w != 16 && p.w != 32) || !p.n2) return -1;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return -1;
  const int N = p.w;
  // the contraction of an inverse pass is never cut for 16 / 32-point kernels (make_inv_plan: only 8-point)
  if (inverse && (p.cutoff1 != N || p.cutoff2 != N)) return -1;
  if (p.shift1 < 1 && inverse) return -1;
  ImmaTr a;
  a.tables = tables(N, p.th, p.tv);
  if (!a.tables) { set_error("transform: table upload failed"); return 1; }
  a.shift1 = p.shift1; a.shift2 = p.shift2;
  if (!inverse) { a.cut_a = p.cutoff1; a.cut_b = p.cutoff2; a.cut_c = p.reduced2; }
  else { a.cut_a = p.reduced1; a.cut_b = N; a.cut_c = N; }
  size_t grid = (n_tus + 3) / 4;
  const size_t cap = (size_t)sm_count() * (N == 32 ? 4 : 10);
  if (grid > cap) grid = cap;
  if (!inverse) {
    if (N == 16) transform_imma_fwd_kernel<16><<<(unsigned)grid, 128, 0, st>>>(in, out, n_tus, a);
    else transform_imma_fwd_kernel<32><<<(unsigned)grid, 128, 0, st>>>(in, out, n_tus, a);
  } else {
    if (N == 16) transform_imma_inv_kernel<16><<<(unsigned)grid, 128, 0, st>>>(in, out, n_tus, a);
    else transform_imma_inv_kernel<32><<<(unsigned)grid, 128, 0, st>>>(in, out, n_tus, a);
  }
  UVGB_LAUNCH_CHECK();
  return 0;
}

}  // namespace uvgb
