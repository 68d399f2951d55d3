// fused_imma.cu - quantize_residual for square 8x8 / 16x16 / 32x32 TUs (8-bit video) on the
// INT8 tensor cores: the whole chain residual -> forward transform -> quant -> dequant ->
// inverse transform -> reconstruction in ONE warp, entirely in registers, in exact integer
// arithmetic (int8 x int8 -> int32 mma.sync.m16n8k16).
//
// Products (same chaining as fused_mma.cu: the C fragment of each product is the A or B
// fragment of the next, so nothing leaves registers between the pixel loads and the stores):
//   F1  D'  = Th * Xref^T + (-Th) * Xpred^T     pixel BYTES are the B operand - no residual ALU
//   F2  Y^T = D' * Tv^T                          A = D' split into bytes (lo u8, hi s8)
//       quant / dequant element-wise on fragments
//   I1  U   = Tv^T * C                           B = dequantised coefficients, byte split
//   I2  R   = U * Th                             A = U, byte split
// A 16-bit value v = hi*256 + lo is multiplied as two byte planes with separate int32
// accumulators recombined as accH*256 + accL: exact by construction (int32 accumulate, all
// |sums| < 2^27). Rounding shifts, int16 wraps and clips of dct-generic.c run on the int32
// accumulators between products.
//
// A thread's C fragment holds columns (2t, 2t+1) of each 8-wide tile, its A/B byte fragment
// needs four consecutive contraction indices: the contraction index is therefore permuted
// (slot 4t+e <-> column 8*(e>>1) + 2t + (e&1) of a 16-wide group) and the constant matrix
// fragments are built with the same permutation. The output-column permutation of I2 makes
// every thread reconstruct 4 CONTIGUOUS pixels of the rows it loaded for F1.
//
// 8x8 TUs: two TUs per warp, stacked in the M dimension with block-diagonal constant operands
// (rows g <-> TU A, rows g+8 <-> TU B), so fragments are fully used.
#include "quant.cuh"
#include "imma.cuh"
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace uvgb {

struct ImmaParams {
  const uint32_t *tables;         // device: [AH | ANH | AVT | BV | BH] fragment words for (N, th, tv)
  int shift_f1, shift_f2, shift_i1, shift_i2;
  int cut1, cut2;                 // forward zero-out: frequencies >= cut are zero (hor, ver)
};

template <int N> struct ImmaShape {
  static constexpr int MT = N >= 16 ? N / 16 : 1;
  static constexpr int NT = N / 8;
  static constexpr int KT = N >= 16 ? N / 16 : 1;
  static constexpr int TUS = N >= 16 ? 1 : 2;        // TUs per warp
};


// Constant operand fragments for (N, th, tv), in the register/lane order mma.sync expects:
//   AH  [mt][kt][r][lane]  F1 A =  Th   (natural contraction order = pixel columns)
//   ANH                    F1 A = -Th
//   AVT [mt][kt][r][lane]  I1 A = Tv^T  (contraction permuted by pi)
//   BV  [kt][nt][lane]     F2 B = Tv^T  (contraction permuted by pi)
//   BH  [kt][nt][lane]     I2 B = Th    (contraction pi, output columns sigma)
// 8x8: two TUs per warp, block-diagonal A operands (row g <-> TU A, row g+8 <-> TU B).
template <int N>
static void build_tables(const int8_t *th, const int8_t *tv, std::vector<uint32_t> &out)
{
  using S = ImmaShape<N>;
  constexpr int MT = S::MT, NT = S::NT, KT = S::KT;
  constexpr int NA = MT * KT * 2 * 32, NB = KT * NT * 32;
  out.assign(3 * NA + 2 * NB, 0);
  for (int e0 = 0; e0 < NA; ++e0) {
    const int ln = e0 & 31, r = (e0 >> 5) & 1, tile = e0 >> 6, kt = tile % KT, mt = tile / KT;
    const int g = ln >> 2, t = ln & 3;
    int bh[4], bv[4];
    if (N >= 16) {
      const int row = mt * 16 + g + 8 * r;
      for (int e = 0; e < 4; ++e) {
        bh[e] = th[row * N + kt * 16 + 4 * t + e];                       // Th[k=row][i]
        bv[e] = tv[(kt * 16 + pi_col(t, e)) * N + row];                  // Tv[kv][iv=row]
      }
    } else {
      for (int e = 0; e < 4; ++e) {
        const int ks = 4 * t + e;                                        // F1 slot: K 0-7 TU A pixels, 8-15 TU B pixels
        bh[e] = ((ks >> 3) == r) ? th[g * N + (ks & 7)] : 0;
        bv[e] = ((e >> 1) == r) ? tv[(2 * t + (e & 1)) * N + g] : 0;    // I1 slot e<2 -> TU A, e>=2 -> TU B
      }
    }
    out[e0] = pack4(bh[0], bh[1], bh[2], bh[3]);
    out[NA + e0] = pack4(-bh[0], -bh[1], -bh[2], -bh[3]);
    out[2 * NA + e0] = pack4(bv[0], bv[1], bv[2], bv[3]);
  }
  for (int e0 = 0; e0 < NB; ++e0) {
    const int ln = e0 & 31, tile = e0 >> 5, nt = tile % NT, kt = tile / NT;
    const int g = ln >> 2, t = ln & 3;
    int bv[4], bh[4];
    if (N >= 16) {
      const int kv = nt * 8 + g;
      const int ih = 16 * (nt >> 1) + 4 * (g >> 1) + 2 * (nt & 1) + (g & 1);   // sigma
      for (int e = 0; e < 4; ++e) {
        const int kk = kt * 16 + pi_col(t, e);
        bv[e] = tv[kv * N + kk];                                         // B[j][kv] = Tv[kv][j]
        bh[e] = th[kk * N + ih];                                         // B[kh][ih] = Th[kh][ih]
      }
    } else {
      for (int e = 0; e < 4; ++e) {
        bv[e] = e < 2 ? tv[g * N + 2 * t + e] : 0;                       // F2: A slots e>=2 are zero padding
        bh[e] = e < 2 ? th[(2 * t + e) * N + g] : 0;                     // I2
      }
    }
    out[3 * NA + e0] = pack4(bv[0], bv[1], bv[2], bv[3]);
    out[3 * NA + NB + e0] = pack4(bh[0], bh[1], bh[2], bh[3]);
  }
}

static const uint32_t *cached_tables(int N, int th, int tv)
{
  static std::mutex mu;
  static const uint32_t *cache[3][3][3] = {};       // [size idx][th][tv]
  const int si = N == 8 ? 0 : N == 16 ? 1 : 2;
  std::lock_guard<std::mutex> lk(mu);
  if (cache[si][th][tv]) return cache[si][th][tv];
  std::vector<uint32_t> h;
  const int8_t *mh = tr_matrix_host(th, N), *mv = tr_matrix_host(tv, N);
  if (N == 8) build_tables<8>(mh, mv, h); else if (N == 16) build_tables<16>(mh, mv, h); else build_tables<32>(mh, mv, h);
  uint32_t *d = nullptr;
  if (cudaMalloc(&d, h.size() * 4) != cudaSuccess) return nullptr;
  if (cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  cache[si][th][tv] = d;
  return d;
}

// PLAIN: no forward zero-out (cut1 == cut2 == N: every DCT-2 pair and every 8 / 16-point MTS pair), so the per-value
// frequency compares disappear. The kernel only ever runs on 8-bit video, so the four rounding shifts are literals and
// their offsets ride in as the initial value of the (low-plane) accumulators.
template <int N, bool PLAIN>
__global__ void __launch_bounds__(128, N == 32 ? 4 : 1)
fused_imma_kernel(FusedArgs a, size_t n_tus, QParams q, ImmaParams p)
{
  using S = ImmaShape<N>;
  constexpr int MT = S::MT, NT = S::NT, KT = S::KT;
  constexpr int L2N = N == 8 ? 3 : N == 16 ? 4 : 5;
  constexpr int SF1 = L2N - 1, SF2 = L2N + 6, SI1 = 7, SI2 = 12;             // dct-generic.c:2613-2615, :2673-2674 at 8 bits
  constexpr int RF1 = SF1 > 0 ? 1 << (SF1 - 1) : 0, RF2 = 1 << (SF2 - 1), RI1 = 1 << (SI1 - 1), RI2 = 1 << (SI2 - 1);
  // constant operand fragments, [tile][reg][lane]
  __shared__ uint32_t sAH[MT * KT * 2 * 32];     // F1  A =  Th        (natural contraction order: pixel columns)
  __shared__ uint32_t sANH[MT * KT * 2 * 32];    // F1  A = -Th
  __shared__ uint32_t sBV[KT * NT * 32];         // F2  B = Tv^T       (contraction permuted by pi)
  __shared__ uint32_t sAVT[MT * KT * 2 * 32];    // I1  A = Tv^T       (contraction permuted by pi)
  __shared__ uint32_t sBH[KT * NT * 32];         // I2  B = Th         (contraction pi, output columns sigma)

  // fragment tables are built once per (N, type pair) on the host (build_tables below) and
  // cached in device memory: a CTA only copies its 1-5 KB image
  {
    constexpr int NA = MT * KT * 2 * 32, NB = KT * NT * 32;
    for (int e0 = threadIdx.x; e0 < NA; e0 += blockDim.x) {
      sAH[e0] = p.tables[e0]; sANH[e0] = p.tables[NA + e0]; sAVT[e0] = p.tables[2 * NA + e0];
    }
    for (int e0 = threadIdx.x; e0 < NB; e0 += blockDim.x) {
      sBV[e0] = p.tables[3 * NA + e0]; sBH[e0] = p.tables[3 * NA + NB + e0];
    }
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const size_t warps_total = (size_t)gridDim.x * (blockDim.x >> 5);
  const size_t n_units = (n_tus + S::TUS - 1) / S::TUS;

  for (size_t unit = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); unit < n_units; unit += warps_total) {
    // ---- pixel loads -------------------------------------------------------------------
    // N >= 16: row tile nt (rows nt*8+g), column group kt: pixels kt*16+4t..+3 of TU `unit`
    // N == 8 : TU 2*unit + (t>>1), row g, pixels 4*(t&1)..+3
    uint32_t refw[NT][KT], predw[NT][KT];
    const uint8_t *pref, *ppred; uint8_t *prec;
    size_t tu0 = unit * S::TUS;
    bool tu_valid = true;
    if (N >= 16) {
      pref = static_cast<const uint8_t *>(a.ref.base) + a.ref.offset(tu0);
      ppred = static_cast<const uint8_t *>(a.pred.base) + a.pred.offset(tu0);
      prec = const_cast<uint8_t *>(static_cast<const uint8_t *>(a.rec.base)) + a.rec.offset(tu0);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
          refw[nt][kt] = __ldg(reinterpret_cast<const uint32_t *>(pref + (size_t)(nt * 8 + g) * a.ref.stride + kt * 16 + 4 * t));
          predw[nt][kt] = __ldg(reinterpret_cast<const uint32_t *>(ppred + (size_t)(nt * 8 + g) * a.pred.stride + kt * 16 + 4 * t));
        }
    } else {
      const size_t mytu = tu0 + (t >> 1);
      tu_valid = mytu < n_tus;
      const size_t ltu = tu_valid ? mytu : tu0;
      pref = static_cast<const uint8_t *>(a.ref.base) + a.ref.offset(ltu);
      ppred = static_cast<const uint8_t *>(a.pred.base) + a.pred.offset(ltu);
      prec = const_cast<uint8_t *>(static_cast<const uint8_t *>(a.rec.base)) + a.rec.offset(ltu);
      refw[0][0] = __ldg(reinterpret_cast<const uint32_t *>(pref + (size_t)g * a.ref.stride + 4 * (t & 1)));
      predw[0][0] = __ldg(reinterpret_cast<const uint32_t *>(ppred + (size_t)g * a.pred.stride + 4 * (t & 1)));
      if (!tu_valid) predw[0][0] = refw[0][0];        // odd tail: the phantom TU has zero residual
    }

    // ---- F1 ----------------------------------------------------------------------------
    int v[MT][NT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        int acc[4] = { RF1, RF1, RF1, RF1 };
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
          const int fi = ((mt * KT + kt) * 2) * 32 + lane;
          imma<false, true>(acc, sAH[fi], sAH[fi + 32], refw[nt][kt]);
          imma<false, true>(acc, sANH[fi], sANH[fi + 32], predw[nt][kt]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int k = N >= 16 ? mt * 16 + g + 8 * (c >> 1) : g;
          const int s = acc[c] >> SF1;                     // its int16 wrap = split4 taking the low half below
          v[mt][nt][c] = (PLAIN || k < p.cut1) ? s : 0;
        }
      }

    // ---- F2 ----------------------------------------------------------------------------
    {
      int y[MT][NT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        int accL[NT][4], accH[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) { accL[nt][c] = RF2; accH[nt][c] = 0; }
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
          uint32_t lo0, hi0, lo1, hi1;
          if (N >= 16) {
            split4(v[mt][2 * kt][0], v[mt][2 * kt][1], v[mt][2 * kt + 1][0], v[mt][2 * kt + 1][1], lo0, hi0);   // row g
            split4(v[mt][2 * kt][2], v[mt][2 * kt][3], v[mt][2 * kt + 1][2], v[mt][2 * kt + 1][3], lo1, hi1);   // row g+8
          } else {
            split4(v[0][0][0], v[0][0][1], 0, 0, lo0, hi0);
            split4(v[0][0][2], v[0][0][3], 0, 0, lo1, hi1);
          }
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const uint32_t b0 = sBV[(kt * NT + nt) * 32 + lane];
            imma<true, false>(accL[nt], lo0, lo1, b0);
            imma<false, false>(accH[nt], hi0, hi1, b0);
          }
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int kv = nt * 8 + 2 * t + (c & 1);
            const int s = (int)(short)((accH[nt][c] * 256 + accL[nt][c]) >> SF2);
            y[mt][nt][c] = (PLAIN || kv < p.cut2) ? s : 0;
          }
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) v[mt][nt][c] = y[mt][nt][c];
    }

    // ---- quant, coefficients out, cost, dequant ------------------------------------------
    // N >= 16: one TU; N == 8: c0,c1 belong to TU A (rows g), c2,c3 to TU B (rows g+8)
    unsigned nzA = 0, nzB = 0, costA = 0, costB = 0;
    const uint32_t wlo = a.w0 | (a.w1 << 16), whi = a.w2 | (a.w3 << 16);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int cf = v[mt][nt][c];
          // no int16 clip of the level: 8-bit video has q_bits >= 16 (checked by the launcher), so
          // |level| <= (32768 * 26214 + add) >> 16 < 13200
          const unsigned ulevel = ((unsigned)abs(cf) * (unsigned)q.coeff_default + (unsigned)q.add) >> q.q_bits;
          const int level = cf < 0 ? -(int)ulevel : (int)ulevel;
          const int kv = nt * 8 + 2 * t + (c & 1);
          // fast_coeff_cost weight of min(|level|, 3): one PRMT on the packed Q8.8 table, no branch
          // (bytes 2, 3 of the PRMT result are copies of byte 0 of the table: a constant, taken off after the reduction)
          const unsigned wgt = __byte_perm(wlo, whi, min(ulevel, 3u) * 0x22u + 0x10u);
          if (N >= 16) {
            const int kh = mt * 16 + g + 8 * (c >> 1);
            a.coeff_out[tu0 * (size_t)(N * N) + kv * N + kh] = (int16_t)level;
            nzA |= ulevel; costA += wgt;
          } else {
            const size_t tuc = tu0 + (c >> 1);
            if (tuc < n_tus) a.coeff_out[tuc * (size_t)(N * N) + kv * N + g] = (int16_t)level;
            if (c >> 1) { nzB |= ulevel; costB += wgt; } else { nzA |= ulevel; costA += wgt; }
          }
          v[mt][nt][c] = (level * q.dq_scale + q.dq_add) >> q.dq_shift;   // dequant; its clip to int16 happens in split4_sat
        }
    const bool hasA = __any_sync(0xffffffffu, nzA != 0);
    const bool hasB = N >= 16 ? false : __any_sync(0xffffffffu, nzB != 0);
    if (a.cost) {
      // every lane added (N * N / 32) [two 8x8 TUs: 2 each] weights per TU, each carrying the constant junk halves
      const unsigned junk = ((wlo & 0xffu) * 0x01010000u) * (N >= 16 ? (unsigned)(N * N) : 64u);
      costA = __reduce_add_sync(0xffffffffu, costA) - junk;
      if (N < 16) costB = __reduce_add_sync(0xffffffffu, costB) - junk;
      if (lane == 0) {
        a.cost[tu0] = (costA + 128u) >> 8;
        if (N < 16 && tu0 + 1 < n_tus) a.cost[tu0 + 1] = (costB + 128u) >> 8;
      }
    }
    if (lane == 0) {
      a.has_coeffs[tu0] = hasA ? 1 : 0;
      if (N < 16 && tu0 + 1 < n_tus) a.has_coeffs[tu0 + 1] = hasB ? 1 : 0;
    }
    const bool any_recon = (hasA || hasB) && !a.early_skip;

    // ---- I1 + I2 (skipped when no TU of the warp has coefficients) --------------------------
    int rr[MT][NT][4];
    if (any_recon) {
      int u[MT][NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        int accL[MT][4], accH[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int c = 0; c < 4; ++c) { accL[mt][c] = RI1; accH[mt][c] = 0; }
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
          uint32_t lo, hi;
          if (N >= 16) {
            const int h2 = 2 * (nt & 1);
            split4_sat(v[nt >> 1][2 * kt][h2], v[nt >> 1][2 * kt][h2 + 1], v[nt >> 1][2 * kt + 1][h2], v[nt >> 1][2 * kt + 1][h2 + 1], lo, hi);
          } else {
            split4_sat(v[0][0][0], v[0][0][1], v[0][0][2], v[0][0][3], lo, hi);
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const int fi = ((mt * KT + kt) * 2) * 32 + lane;
            imma<false, true>(accL[mt], sAVT[fi], sAVT[fi + 32], lo);
            imma<false, false>(accH[mt], sAVT[fi], sAVT[fi + 32], hi);
          }
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int c = 0; c < 4; ++c)
            u[mt][nt][c] = (accH[mt][c] * 256 + accL[mt][c]) >> SI1;       // clipped to int16 by split4_sat below
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        int accL[NT][4], accH[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) { accL[nt][c] = RI2; accH[nt][c] = 0; }
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
          uint32_t lo0, hi0, lo1, hi1;
          if (N >= 16) {
            split4_sat(u[mt][2 * kt][0], u[mt][2 * kt][1], u[mt][2 * kt + 1][0], u[mt][2 * kt + 1][1], lo0, hi0);
            split4_sat(u[mt][2 * kt][2], u[mt][2 * kt][3], u[mt][2 * kt + 1][2], u[mt][2 * kt + 1][3], lo1, hi1);
          } else {
            split4_sat(u[0][0][0], u[0][0][1], 0, 0, lo0, hi0);
            split4_sat(u[0][0][2], u[0][0][3], 0, 0, lo1, hi1);
          }
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const uint32_t b0 = sBH[(kt * NT + nt) * 32 + lane];
            imma<true, false>(accL[nt], lo0, lo1, b0);
            imma<false, false>(accH[nt], hi0, hi1, b0);
          }
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c)
            rr[mt][nt][c] = min((accH[nt][c] * 256 + accL[nt][c]) >> SI2, 32767);   // (the lower int16 bound cannot matter: below)
      }
    }

    // ---- reconstruction + SSD ------------------------------------------------------------
    if (N >= 16) {
      const bool recon = any_recon;
      unsigned ssd = 0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)          // row tile of the pixel words
#pragma unroll
        for (int qs = 0; qs < KT; ++qs) {
          const int mt = nt >> 1, hh = nt & 1;
          uint32_t outw = predw[nt][qs];
          if (recon) {
            // reference: clip(0, 255, (int16)(clip16(r) + pred)). r <= 32767 is applied above; r < -32768 gives a negative
            // sum with or without its clip, which the saturating pack turns into 0 either way; the int16 wrap of the sum
            // (r close to 32767) is kept
            int px[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int r = rr[mt][2 * qs + (e >> 1)][2 * hh + (e & 1)];
              px[e] = (int)(short)(max(r, -32768) + (int)((predw[nt][qs] >> (8 * e)) & 0xffu));
            }
            outw = pack4_sat_u8(px[0], px[1], px[2], px[3]);
          }
          if (recon || !a.rec_is_pred)
            *reinterpret_cast<uint32_t *>(prec + (size_t)(nt * 8 + g) * a.rec.stride + qs * 16 + 4 * t) = outw;
          const uint32_t d4 = __vabsdiffu4(refw[nt][qs], outw);
          ssd = __dp4a(d4, d4, ssd);
        }
      if (a.ssd) {
        ssd = __reduce_add_sync(0xffffffffu, ssd);
        if (lane == 0) a.ssd[tu0] = (uint32_t)((int32_t)ssd >> q.ssd_shift);
      }
    } else {
      // thread holds (TU A, row g, px 2t,2t+1) in rr[..][0],[1] and (TU B, row g, px 2t,2t+1) in [2],[3];
      // pred/ref words live in lane (g, TU*2 + (t>>1)): px 4*(t>>1).. -> fetch and pick the half
      const uint32_t pA = __shfl_sync(0xffffffffu, predw[0][0], (g << 2) | (t >> 1));
      const uint32_t pB = __shfl_sync(0xffffffffu, predw[0][0], (g << 2) | 2 | (t >> 1));
      const uint32_t rA = __shfl_sync(0xffffffffu, refw[0][0], (g << 2) | (t >> 1));
      const uint32_t rB = __shfl_sync(0xffffffffu, refw[0][0], (g << 2) | 2 | (t >> 1));
      const int sh = 16 * (t & 1);
      unsigned ssdA = 0, ssdB = 0;
#pragma unroll
      for (int tb = 0; tb < 2; ++tb) {
        const bool recon = (tb ? hasB : hasA) && !a.early_skip;
        const uint32_t pw = ((tb ? pB : pA) >> sh) & 0xffffu, rw = ((tb ? rB : rA) >> sh) & 0xffffu;
        uint32_t ow = pw;
        if (recon) {
          const int x0 = clip3(0, q.pixel_max, (int)(short)(max(rr[0][0][2 * tb], -32768) + (int)(pw & 0xffu)));
          const int x1 = clip3(0, q.pixel_max, (int)(short)(max(rr[0][0][2 * tb + 1], -32768) + (int)(pw >> 8)));
          ow = (uint32_t)x0 | ((uint32_t)x1 << 8);
        }
        const size_t tuc = tu0 + tb;
        if (tuc < n_tus && (recon || !a.rec_is_pred)) {
          uint8_t *dst = const_cast<uint8_t *>(static_cast<const uint8_t *>(a.rec.base)) + a.rec.offset(tuc) +
                         (size_t)g * a.rec.stride + 2 * t;
          *reinterpret_cast<uint16_t *>(dst) = (uint16_t)ow;
        }
        const uint32_t d4 = __vabsdiffu4(rw, ow);
        if (tb) ssdB = __dp4a(d4, d4, ssdB); else ssdA = __dp4a(d4, d4, ssdA);
      }
      if (a.ssd) {
        ssdA = __reduce_add_sync(0xffffffffu, ssdA); ssdB = __reduce_add_sync(0xffffffffu, ssdB);
        if (lane == 0) {
          a.ssd[tu0] = (uint32_t)((int32_t)ssdA >> q.ssd_shift);
          if (tu0 + 1 < n_tus) a.ssd[tu0 + 1] = (uint32_t)((int32_t)ssdB >> q.ssd_shift);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// 8x8 specialisation (the kernel of the path with the lowest bytes/s: 120 TUs per CTU).
// Same four products and the same two-TUs-per-warp stacking as fused_imma_kernel<8>, with the
// instruction stream cut to what an 8-bit 8x8 TU needs:
//   * a lane loads the two pixels (2t, 2t+1) of row g of BOTH TUs (four 16-bit loads); the F1
//     contraction slots are permuted to match (slot 4t+e <-> TU e>>1, pixel 2t+(e&1)), so the
//     inverse transform hands every lane back exactly the pixels it loaded: no shuffles;
//   * the eight constant fragment words live in registers for the whole grid-stride loop (no
//     shared memory, no barrier);
//   * the rounding offsets enter as the C operand of the first mma of each stage;
//   * shifts are literals (8-bit video: 2 / 9 / 7 / 12), there is no zero-out for 8-point
//     transforms (skip_w/h only exist for 32-point, dct-generic.c:2582-2583), and the level
//     clip is dropped: q_bits >= 18 here (14 + qp/6 + transform_shift 4, quant-generic.c:62-77),
//     so |level| <= (32768 * 26214 + add) >> 18 < 3300;
//   * fast_coeff_cost weights are picked with one PRMT from the packed Q8.8 table instead of a
//     4-way branch; 32-bit block indexing.
template <bool AU, bool BU>
__device__ __forceinline__ void imma_c(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0, int c)
{
  if (AU && !BU)
    asm("mma.sync.aligned.m16n8k16.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%7,%7,%7};\n"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(a0), "r"(a1), "r"(b0), "r"(c));
  else if (!AU && BU)
    asm("mma.sync.aligned.m16n8k16.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%7,%7,%7};\n"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(a0), "r"(a1), "r"(b0), "r"(c));
  else
    asm("mma.sync.aligned.m16n8k16.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%7,%7,%7};\n"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(a0), "r"(a1), "r"(b0), "r"(c));
}

// byte offset of block i (i, pitches < 2^32: checked on the host)
__device__ __forceinline__ uint64_t block_off32(const Blocks &b, uint32_t i)
{
  if (b.bpr == 1) return (uint64_t)i * (uint32_t)b.row_pitch;
  const uint32_t r = i / b.bpr, c = i - r * b.bpr;
  return (uint64_t)r * (uint32_t)b.row_pitch + (uint64_t)c * (uint32_t)b.col_pitch;
}

// fragment words of lane (g, t), [word][lane]: AH0 AH1 ANH0 ANH1 AVT0 AVT1 BV BH
static void build_tables8(const int8_t *th, const int8_t *tv, std::vector<uint32_t> &out)
{
  out.assign(8 * 32, 0);
  for (int ln = 0; ln < 32; ++ln) {
    const int g = ln >> 2, t = ln & 3;
    for (int r = 0; r < 2; ++r) {
      int bh[4], bv[4];
      for (int e = 0; e < 4; ++e) {
        bh[e] = ((e >> 1) == r) ? th[g * 8 + 2 * t + (e & 1)] : 0;      // F1: slot e -> TU e>>1, pixel 2t+(e&1)
        bv[e] = ((e >> 1) == r) ? tv[(2 * t + (e & 1)) * 8 + g] : 0;    // I1: slot e -> TU e>>1, kv 2t+(e&1)
      }
      out[(0 + r) * 32 + ln] = pack4(bh[0], bh[1], bh[2], bh[3]);
      out[(2 + r) * 32 + ln] = pack4(-bh[0], -bh[1], -bh[2], -bh[3]);
      out[(4 + r) * 32 + ln] = pack4(bv[0], bv[1], bv[2], bv[3]);
    }
    out[6 * 32 + ln] = pack4(tv[g * 8 + 2 * t], tv[g * 8 + 2 * t + 1], 0, 0);             // F2 B: slots e>=2 are padding
    out[7 * 32 + ln] = pack4(th[(2 * t) * 8 + g], th[(2 * t + 1) * 8 + g], 0, 0);         // I2 B
  }
}

static const uint32_t *cached_tables8(int th, int tv)
{
  static std::mutex mu;
  static const uint32_t *cache[3][3] = {};
  std::lock_guard<std::mutex> lk(mu);
  if (cache[th][tv]) return cache[th][tv];
  std::vector<uint32_t> h;
  build_tables8(tr_matrix_host(th, 8), tr_matrix_host(tv, 8), h);
  uint32_t *d = nullptr;
  if (cudaMalloc(&d, h.size() * 4) != cudaSuccess) return nullptr;
  if (cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  cache[th][tv] = d;
  return d;
}

// DENSE: ref, pred and rec are dense arrays of 8x8 blocks (stride 8, 64 bytes per TU) - the layout of every batched
// caller; block addresses are then tu * 64 and the descriptor arithmetic (and its bpr == 1 branches) disappears.
template <bool DENSE>
__global__ void __launch_bounds__(128)
fused_imma8_kernel(FusedArgs a, uint32_t n_tus, QParams q, const uint32_t *__restrict__ tab)
{
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const uint32_t ah0 = tab[lane], ah1 = tab[32 + lane], anh0 = tab[64 + lane], anh1 = tab[96 + lane];
  const uint32_t avt0 = tab[128 + lane], avt1 = tab[160 + lane], bv = tab[192 + lane], bh = tab[224 + lane];
  const uint32_t wlo = a.w0 | (a.w1 << 16), whi = a.w2 | (a.w3 << 16);
  const uint32_t n_units = (n_tus + 1) >> 1;
  const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
  const uint8_t *ref_b = static_cast<const uint8_t *>(a.ref.base), *pred_b = static_cast<const uint8_t *>(a.pred.base);
  uint8_t *rec_b = const_cast<uint8_t *>(static_cast<const uint8_t *>(a.rec.base));
  const uint32_t ref_ln = g * (DENSE ? 8u : a.ref.stride) + 2 * t, pred_ln = g * (DENSE ? 8u : a.pred.stride) + 2 * t,
                 rec_ln = g * (DENSE ? 8u : a.rec.stride) + 2 * t;
  auto off_ref = [&](uint32_t tu) -> uint64_t { return DENSE ? (uint64_t)tu * 64u : block_off32(a.ref, tu); };
  auto off_pred = [&](uint32_t tu) -> uint64_t { return DENSE ? (uint64_t)tu * 64u : block_off32(a.pred, tu); };
  auto off_rec = [&](uint32_t tu) -> uint64_t { return DENSE ? (uint64_t)tu * 64u : block_off32(a.rec, tu); };
  const bool do_recon = !a.early_skip;

  for (uint32_t unit = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); unit < n_units; unit += warps_total) {
    const uint32_t tuA = 2 * unit;
    const bool validB = tuA + 1 < n_tus;
    const uint32_t tuB = validB ? tuA + 1 : tuA;
    // ---- pixels (2t, 2t+1) of row g of both TUs -----------------------------------------------
    const uint32_t rA = __ldg(reinterpret_cast<const uint16_t *>(ref_b + off_ref(tuA) + ref_ln));
    const uint32_t rB = __ldg(reinterpret_cast<const uint16_t *>(ref_b + off_ref(tuB) + ref_ln));
    const uint32_t pA = __ldg(reinterpret_cast<const uint16_t *>(pred_b + off_pred(tuA) + pred_ln));
    uint32_t pB = __ldg(reinterpret_cast<const uint16_t *>(pred_b + off_pred(tuB) + pred_ln));
    if (!validB) pB = rB;                                  // odd tail: the phantom TU has zero residual
    const uint32_t refw = rA | (rB << 16), predw = pA | (pB << 16);

    // ---- F1: D' = Th X^T (rows kh of TU A | TU B, columns = pixel rows 2t, 2t+1) ----------------
    int v[4];
    {
      int d[4];
      imma_c<false, true>(d, ah0, ah1, refw, 2);
      imma<false, true>(d, anh0, anh1, predw);
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = d[c] >> 2;            // (int16 wrap = the byte picks of the next operands)
    }
    // ---- F2: Y^T = D' Tv^T (rows kh, columns kv = 2t, 2t+1); bytes 2,3 of the A words meet zeros --
    {
      int accL[4], accH[4];
      imma_c<true, false>(accL, __byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), bv, 256);
      imma_c<false, false>(accH, __byte_perm(v[0], v[1], 0x0051), __byte_perm(v[2], v[3], 0x0051), bv, 0);
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = (int)(short)((accH[c] * 256 + accL[c]) >> 9);
    }
    // ---- quant, cost, dequant: c0,c1 = TU A (kh g, kv 2t, 2t+1), c2,c3 = TU B --------------------
    unsigned ul[4], wg[4];
    int lv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      ul[c] = ((unsigned)abs(v[c]) * (unsigned)q.coeff_default + (unsigned)q.add) >> q.q_bits;
      lv[c] = v[c] < 0 ? -(int)ul[c] : (int)ul[c];
      wg[c] = __byte_perm(wlo, whi, min(ul[c], 3u) * 0x22u + 0x10u);        // + constant junk in bytes 2, 3: removed below
      v[c] = (lv[c] * q.dq_scale + q.dq_add) >> q.dq_shift;     // clipped to int16 by split4_sat below
    }
    {
      int16_t *co = a.coeff_out + (size_t)tuA * 64 + 16 * t + g;           // [kv][kh]
      co[0] = (int16_t)lv[0]; co[8] = (int16_t)lv[1];
      if (validB) { co[64] = (int16_t)lv[2]; co[72] = (int16_t)lv[3]; }
    }
    const bool hasA = __any_sync(0xffffffffu, (ul[0] | ul[1]) != 0);
    const bool hasB = __any_sync(0xffffffffu, (ul[2] | ul[3]) != 0);
    if (a.cost) {
      const unsigned junk = ((wlo & 0xffu) * 0x01010000u) * 64u;           // 64 weights per TU, each with the junk halves
      const unsigned costA = __reduce_add_sync(0xffffffffu, wg[0] + wg[1]) - junk;
      const unsigned costB = __reduce_add_sync(0xffffffffu, wg[2] + wg[3]) - junk;
      if (lane == 0) {
        a.cost[tuA] = (costA + 128u) >> 8;
        if (validB) a.cost[tuA + 1] = (costB + 128u) >> 8;
      }
    }
    if (lane == 0) {
      a.has_coeffs[tuA] = hasA ? 1 : 0;
      if (validB) a.has_coeffs[tuA + 1] = hasB ? 1 : 0;
    }
    const bool reconA = hasA && do_recon, reconB = hasB && do_recon;

    // ---- I1, I2, reconstruction (skipped when neither TU has coefficients) ---------------------
    uint32_t outw = predw;
    if (reconA || reconB) {
      uint32_t lo, hi;
      split4_sat(v[0], v[1], v[2], v[3], lo, hi);
      int accL[4], accH[4];
      imma_c<false, true>(accL, avt0, avt1, lo, 64);
      imma_c<false, false>(accH, avt0, avt1, hi, 0);
      // U clipped to int16 and packed in pairs by one saturating pack each; byte planes (u0.b0, u1.b0, 0, 0) / (.b1 ..)
      const uint32_t u01 = pack2_sat_s16((accH[0] * 256 + accL[0]) >> 7, (accH[1] * 256 + accL[1]) >> 7);
      const uint32_t u23 = pack2_sat_s16((accH[2] * 256 + accL[2]) >> 7, (accH[3] * 256 + accL[3]) >> 7);
      imma_c<true, false>(accL, __byte_perm(u01, 0, 0x4420), __byte_perm(u23, 0, 0x4420), bh, 2048);
      imma_c<false, false>(accH, __byte_perm(u01, 0, 0x4431), __byte_perm(u23, 0, 0x4431), bh, 0);
      // reference: clip(0, 255, (int16)(clip16(r) + pred)): see fused_imma_kernel
      int px[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int r = min(max((accH[c] * 256 + accL[c]) >> 12, -32768), 32767);
        px[c] = (int)(short)(r + (int)((predw >> (8 * c)) & 0xffu));
      }
      const uint32_t w = pack4_sat_u8(px[0], px[1], px[2], px[3]);
      outw = (reconA ? (w & 0xffffu) : (predw & 0xffffu)) | (reconB ? (w & 0xffff0000u) : (predw & 0xffff0000u));
    }
    if (reconA || !a.rec_is_pred)
      *reinterpret_cast<uint16_t *>(rec_b + off_rec(tuA) + rec_ln) = (uint16_t)outw;
    if (validB && (reconB || !a.rec_is_pred))
      *reinterpret_cast<uint16_t *>(rec_b + off_rec(tuB) + rec_ln) = (uint16_t)(outw >> 16);
    if (a.ssd) {
      const uint32_t d4 = __vabsdiffu4(refw, outw);
      const unsigned ssdA = __reduce_add_sync(0xffffffffu, __dp4a(d4 & 0xffffu, d4, 0u));
      const unsigned ssdB = __reduce_add_sync(0xffffffffu, __dp4a(d4 & 0xffff0000u, d4, 0u));
      if (lane == 0) {
        a.ssd[tuA] = (uint32_t)((int32_t)ssdA >> q.ssd_shift);
        if (validB) a.ssd[tuA + 1] = (uint32_t)((int32_t)ssdB >> q.ssd_shift);
      }
    }
  }
}

static bool fits32(const Blocks &b) { return b.row_pitch < (1ull << 32) && b.col_pitch < (1ull << 32); }

// -1 = not covered (the general fused_imma_kernel<8> takes it)
static int launch_imma8(const FusedArgs &a, size_t n_tus, const QParams &q, int th, int tv, cudaStream_t st)
{
  static const char *force = getenv("UVG266_B200_FUSED8");
  if (force && !strcmp(force, "general")) return -1;
  if (n_tus >= (1ull << 31) || q.q_bits < 16 || q.q_bits > 31 || !fits32(a.ref) || !fits32(a.pred) || !fits32(a.rec)) return -1;
  const uint32_t *tab = cached_tables8(th, tv);
  if (!tab) { set_error("fused_imma8: fragment table allocation failed"); return 1; }
  const int warps_per_block = 4;
  const size_t n_units = (n_tus + 1) / 2;
  size_t grid = (n_units + warps_per_block - 1) / warps_per_block;
  const size_t cap = (size_t)sm_count() * 16;
  if (grid > cap) grid = cap;
  auto dense8 = [](const Blocks &b) { return b.bpr == 1 && b.stride == 8 && b.row_pitch == 64; };
  if (dense8(a.ref) && dense8(a.pred) && dense8(a.rec)) fused_imma8_kernel<true><<<(unsigned)grid, warps_per_block * 32, 0, st>>>(a, (uint32_t)n_tus, q, tab);
  else fused_imma8_kernel<false><<<(unsigned)grid, warps_per_block * 32, 0, st>>>(a, (uint32_t)n_tus, q, tab);
  UVGB_LAUNCH_CHECK();
  return 0;
}

template <int N>
static int launch_imma(const FusedArgs &a, size_t n_tus, const QParams &q, const ImmaParams &p, cudaStream_t st)
{
  const int warps_per_block = 4;
  const size_t n_units = (n_tus + ImmaShape<N>::TUS - 1) / ImmaShape<N>::TUS;
  size_t grid = (n_units + warps_per_block - 1) / warps_per_block;
  static const int cap_mult = getenv("UVG266_B200_IMMA_CAP") ? atoi(getenv("UVG266_B200_IMMA_CAP")) : 16;
  const size_t cap = (size_t)sm_count() * (size_t)(cap_mult > 0 ? cap_mult : 16);
  if (grid > cap) grid = cap;
  if (p.cut1 == N && p.cut2 == N) fused_imma_kernel<N, true><<<(unsigned)grid, warps_per_block * 32, 0, st>>>(a, n_tus, q, p);
  else fused_imma_kernel<N, false><<<(unsigned)grid, warps_per_block * 32, 0, st>>>(a, n_tus, q, p);
  UVGB_LAUNCH_CHECK();
  return 0;
}

static bool word_aligned(const Blocks &b, int n)
{
  // every row start must be 4-byte aligned (8x8 rec stores need 2-byte alignment only, covered)
  return (reinterpret_cast<uintptr_t>(b.base) & 3) == 0 && (b.stride & 3) == 0 && (b.col_pitch & 3) == 0 &&
         (b.row_pitch & 3) == 0 && n % 4 == 0;
}

// 8-bit video, square 8/16/32, no sign hiding / transform skip, word-aligned rows. -1 = not covered.
int launch_fused_imma(const FusedArgs &a, size_t n_tus, const QParams &q, int bitdepth, int th, int tv, int skip_w,
                      int skip_h, cudaStream_t st)
{
  if (bitdepth != 8 || q.w != q.h || (q.w != 8 && q.w != 16 && q.w != 32)) return -1;
  if (q.signhide || q.lfnst_idx || a.use_trskip) return -1;
  if (q.q_bits < 16 || q.q_bits > 31) return -1;       // the kernels drop the level clip (see the quant step)
  if (!word_aligned(a.ref, q.w) || !word_aligned(a.pred, q.w) || !word_aligned(a.rec, q.w)) return -1;
  const int N = q.w, l2 = ilog2_host(N);
  ImmaParams p;
  p.tables = cached_tables(N, th, tv);
  if (!p.tables) { set_error("fused_imma: fragment table allocation failed"); return 1; }
  p.shift_f1 = l2 - 1 + bitdepth - 8; p.shift_f2 = l2 + 6; p.shift_i1 = 7; p.shift_i2 = 20 - bitdepth;
  p.cut1 = (th == UVG_B200_DCT2 || N < 8) ? N : N - skip_w;
  p.cut2 = (tv == UVG_B200_DCT2 || N < 8) ? N : N - skip_h;
  switch (N) {
    case 8: {
      const int rc = launch_imma8(a, n_tus, q, th, tv, st);
      return rc >= 0 ? rc : launch_imma<8>(a, n_tus, q, p, st);
    }
    case 16: return launch_imma<16>(a, n_tus, q, p, st);
    default: return launch_imma<32>(a, n_tus, q, p, st);
  }
}

}  // namespace uvgb
