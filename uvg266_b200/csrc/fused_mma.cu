// fused_mma.cu - quantize_residual for square 16x16 and 32x32 TUs on the tensor cores.
//
// The separable transforms of a 16/32-point TU are dense N x N x N contractions
// (Y = Tv * X * Th^T with per-pass rounding) - the one place on this path where tensor cores
// are the right tool. One warp owns one TU and chains four products with
// mma.sync.m16n8k16 (f16 operands, f32 accumulators), keeping everything in registers:
//
//   F1  D'  = Th * X^T          A = Th (constant), B = residual straight from the pixel loads
//   F2  Y^T = D' * Tv^T         A = D' (the C fragments of F1 ARE A fragments), B = Tv^T
//       quant / dequant element-wise on the fragments
//   I1  U   = Tv^T * C          A = Tv^T (constant), B = dequantised coefficients (C fragments
//                               of F2 ARE B fragments)
//   I2  R   = U * Th            A = U (C fragments of I1), B = Th (constant)
//       reconstruction / SSD element-wise
//
// Exactness (outputs must equal the integer reference bit for bit): f16 holds integers up to
// 2048 exactly and an f32 accumulator holds integer sums below 2^24 exactly. Residuals
// (|r| <= 1023) and matrix entries (|c| <= 90) fit directly; 16-bit intermediates are split
// v = hi*256 + lo (hi in [-128,127], lo in [0,255]) and multiplied as two products whose
// accumulators (<= 32*255*90 < 2^20) are recombined in int32. Every rounding shift / wrap /
// clip of dct-generic.c happens in integer registers between the products.
//
// A column permutation of the contraction index (pi) makes each thread's B operand of F1, and
// its reconstruction outputs of I2, four CONTIGUOUS pixels, so pixels move as 32/64-bit words.
#include "quant.cuh"
#include <cuda_fp16.h>

namespace uvgb {

struct MmaParams {
  const int8_t *th, *tv;          // device matrices T[k*N+i]
  int shift_f1, shift_f2, shift_i1, shift_i2;
  int cut1, cut2;                 // forward zero-out: frequencies >= cut are zero (hor, ver)
};

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_h2(int x, int y)
{
  const __half2 h = __halves2half2(__int2half_rn(x), __int2half_rn(y));
  return *reinterpret_cast<const uint32_t *>(&h);
}

// 4 consecutive pixels <-> registers
template <typename Pix> struct Px4;
template <> struct Px4<uint8_t> {
  uint32_t w;
  __device__ __forceinline__ void load(const uint8_t *p) {
    if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) w = __ldg(reinterpret_cast<const uint32_t *>(p));
    else w = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16) | ((uint32_t)__ldg(p + 3) << 24);
  }
  __device__ __forceinline__ int at(int i) const { return (int)((w >> (8 * i)) & 0xffu); }
  __device__ __forceinline__ void set(int v0, int v1, int v2, int v3) { w = (uint32_t)v0 | ((uint32_t)v1 << 8) | ((uint32_t)v2 << 16) | ((uint32_t)v3 << 24); }
  __device__ __forceinline__ void store(uint8_t *p) const {
    if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) *reinterpret_cast<uint32_t *>(p) = w;
    else { p[0] = (uint8_t)w; p[1] = (uint8_t)(w >> 8); p[2] = (uint8_t)(w >> 16); p[3] = (uint8_t)(w >> 24); }
  }
};
template <> struct Px4<uint16_t> {
  uint32_t w0, w1;
  __device__ __forceinline__ void load(const uint16_t *p) {
    if ((reinterpret_cast<uintptr_t>(p) & 7) == 0) { const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p)); w0 = v.x; w1 = v.y; }
    else { w0 = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 16); w1 = (uint32_t)__ldg(p + 2) | ((uint32_t)__ldg(p + 3) << 16); }
  }
  __device__ __forceinline__ int at(int i) const { return (int)(((i < 2 ? w0 : w1) >> (16 * (i & 1))) & 0xffffu); }
  __device__ __forceinline__ void set(int v0, int v1, int v2, int v3) { w0 = (uint32_t)v0 | ((uint32_t)v1 << 16); w1 = (uint32_t)v2 | ((uint32_t)v3 << 16); }
  __device__ __forceinline__ void store(uint16_t *p) const {
    if ((reinterpret_cast<uintptr_t>(p) & 7) == 0) *reinterpret_cast<uint2 *>(p) = make_uint2(w0, w1);
    else { p[0] = (uint16_t)w0; p[1] = (uint16_t)(w0 >> 16); p[2] = (uint16_t)w1; p[3] = (uint16_t)(w1 >> 16); }
  }
};

template <typename Pix, int N>
__global__ void __launch_bounds__(128)
fused_mma_kernel(FusedArgs a, size_t n_tus, QParams q, MmaParams p)
{
  constexpr int MT = N / 16, NT = N / 8, KT = N / 16;
  // constant operand fragments, fragment-major: [tile][reg][lane] -> conflict-free LDS
  __shared__ uint32_t sAH[MT * KT * 4 * 32];    // A = Th, contraction index permuted by pi
  __shared__ uint32_t sBV[KT * NT * 2 * 32];    // B = Tv^T
  __shared__ uint32_t sAVT[MT * KT * 4 * 32];   // A = Tv^T
  __shared__ uint32_t sBH[KT * NT * 2 * 32];    // B = Th, output column index permuted by pi

  for (int e = threadIdx.x; e < MT * KT * 4 * 32; e += blockDim.x) {
    const int ln = e & 31, r = (e >> 5) & 3, tile = e >> 7, kt = tile % KT, mt = tile / KT;
    const int g = ln >> 2, t = ln & 3;
    const int row = mt * 16 + g + 8 * (r & 1);
    // A = Th[k=row][i]: slot 2t+e (+8 for r>>1) of k-tile kt -> pixel column kt*16 + 4t + 2*(r>>1) + e
    const int i0 = kt * 16 + 4 * t + 2 * (r >> 1);
    sAH[e] = pack_h2(p.th[row * N + i0], p.th[row * N + i0 + 1]);
    // A = Tv^T[iv=row][kv] = Tv[kv][iv], natural contraction order
    const int k0 = kt * 16 + 2 * t + 8 * (r >> 1);
    sAVT[e] = pack_h2(p.tv[k0 * N + row], p.tv[(k0 + 1) * N + row]);
  }
  for (int e = threadIdx.x; e < KT * NT * 2 * 32; e += blockDim.x) {
    const int ln = e & 31, r = (e >> 5) & 1, tile = e >> 6, nt = tile % NT, kt = tile / NT;
    const int g = ln >> 2, t = ln & 3;
    const int k0 = kt * 16 + 2 * t + 8 * r;
    // B = Tv^T[j=k][kv=n] = Tv[kv][j]
    const int kv = nt * 8 + g;
    sBV[e] = pack_h2(p.tv[kv * N + k0], p.tv[kv * N + k0 + 1]);
    // B = Th[kh=k][ih], output column slot (nt, g) -> pixel column 16*(nt>>1) + 4*(g>>1) + 2*(nt&1) + (g&1)
    const int ih = 16 * (nt >> 1) + 4 * (g >> 1) + 2 * (nt & 1) + (g & 1);
    sBH[e] = pack_h2(p.th[k0 * N + ih], p.th[(k0 + 1) * N + ih]);
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const size_t warps_total = (size_t)gridDim.x * (blockDim.x >> 5);
  const int add_f1 = p.shift_f1 > 0 ? 1 << (p.shift_f1 - 1) : 0;
  const int add_f2 = 1 << (p.shift_f2 - 1), add_i1 = 1 << (p.shift_i1 - 1), add_i2 = 1 << (p.shift_i2 - 1);

  for (size_t tu = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tu < n_tus; tu += warps_total) {
    const Pix *pref = static_cast<const Pix *>(a.ref.base) + a.ref.offset(tu);
    const Pix *ppred = static_cast<const Pix *>(a.pred.base) + a.pred.offset(tu);
    Pix *prec = const_cast<Pix *>(static_cast<const Pix *>(a.rec.base)) + a.rec.offset(tu);

    // pixels: row tile nt (rows nt*8+g), 16-column segment kt, columns kt*16+4t..+3
    Px4<Pix> refw[NT][KT], predw[NT][KT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        refw[nt][kt].load(pref + (size_t)(nt * 8 + g) * a.ref.stride + kt * 16 + 4 * t);
        predw[nt][kt].load(ppred + (size_t)(nt * 8 + g) * a.pred.stride + kt * 16 + 4 * t);
      }

    // ---- F1: D'[k][j] = sum_i Th[k][i] X[j][i] -------------------------------------------
    int v[MT][NT][4];
    {
      float acc[MT][NT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[mt][nt][c] = 0.f;
#pragma unroll
      for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int x0 = refw[nt][kt].at(0) - predw[nt][kt].at(0), x1 = refw[nt][kt].at(1) - predw[nt][kt].at(1);
          const int x2 = refw[nt][kt].at(2) - predw[nt][kt].at(2), x3 = refw[nt][kt].at(3) - predw[nt][kt].at(3);
          const uint32_t b0 = pack_h2(x0, x1), b1 = pack_h2(x2, x3);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            uint32_t af[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) af[r] = sAH[((mt * KT + kt) * 4 + r) * 32 + lane];
            mma16816(acc[mt][nt], af, b0, b1);
          }
        }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int k = mt * 16 + g + 8 * (c >> 1);
            const int s = (int)(short)((__float2int_rn(acc[mt][nt][c]) + add_f1) >> p.shift_f1);
            v[mt][nt][c] = k < p.cut1 ? s : 0;
          }
    }

    // ---- F2: Y^T[kh][kv] = sum_j D'[kh][j] Tv[kv][j] --------------------------------------
    {
      float accH[MT][NT][4], accL[MT][NT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) { accH[mt][nt][c] = 0.f; accL[mt][nt][c] = 0.f; }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
          uint32_t aH[4], aL[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int x = v[mt][2 * kt + (r >> 1)][2 * (r & 1)], y = v[mt][2 * kt + (r >> 1)][2 * (r & 1) + 1];
            aH[r] = pack_h2(x >> 8, y >> 8); aL[r] = pack_h2(x & 255, y & 255);
          }
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const uint32_t b0 = sBV[((kt * NT + nt) * 2 + 0) * 32 + lane], b1 = sBV[((kt * NT + nt) * 2 + 1) * 32 + lane];
            mma16816(accH[mt][nt], aH, b0, b1);
            mma16816(accL[mt][nt], aL, b0, b1);
          }
        }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int kv = nt * 8 + 2 * t + (c & 1);
            const int y = __float2int_rn(accH[mt][nt][c]) * 256 + __float2int_rn(accL[mt][nt][c]);
            const int s = (int)(short)((y + add_f2) >> p.shift_f2);
            v[mt][nt][c] = kv < p.cut2 ? s : 0;         // coefficient (kh = mt*16+g+8*(c>>1), kv)
          }
    }

    // ---- quant, coefficients out, has_coeffs, cost, dequant --------------------------------
    unsigned nz = 0, cost = 0;
    int16_t *cout = a.coeff_out + tu * (size_t)(N * N);
    const uint32_t wlo = a.w0 | (a.w1 << 16), whi = a.w2 | (a.w3 << 16);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int cf = v[mt][nt][c];
          // no int16 clip of the level: the launcher admits q_bits >= 15 only (|level| <= 26214 + 1)
          const unsigned ul = ((unsigned)abs(cf) * (unsigned)q.coeff_default + (unsigned)q.add) >> q.q_bits;
          const int level = cf < 0 ? -(int)ul : (int)ul;
          const int kh = mt * 16 + g + 8 * (c >> 1), kv = nt * 8 + 2 * t + (c & 1);
          cout[kv * N + kh] = (int16_t)level;
          nz |= ul;
          cost += __byte_perm(wlo, whi, min(ul, 3u) * 0x22u + 0x10u) & 0xffffu;   // fast_coeff_cost weight, packed Q8.8 table
          v[mt][nt][c] = clip3(-32768, 32767, (level * q.dq_scale + q.dq_add) >> q.dq_shift);   // dequant
        }
    const bool has = __any_sync(0xffffffffu, nz != 0);
    const bool recon = has && !a.early_skip;
    if (a.cost) {
      cost = (unsigned)warp_sum((int)cost);
      if (lane == 0) a.cost[tu] = (cost + 128u) >> 8;
    }
    if (lane == 0) a.has_coeffs[tu] = has ? 1 : 0;

    unsigned ssd = 0;
    if (recon) {
      // ---- I1: U[iv][kh] = sum_kv Tv[kv][iv] C[kv][kh]; B fragments from v (layout kh x kv) ----
      int u[MT][NT][4];
      {
        float accH[MT][NT][4], accL[MT][NT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int c = 0; c < 4; ++c) { accH[mt][nt][c] = 0.f; accL[mt][nt][c] = 0.f; }
#pragma unroll
        for (int kt = 0; kt < KT; ++kt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            // n = kh = nt*8+g lives in v[nt>>1][.][2*(nt&1) + e]; k = kv = kt*16 + 2t(+8) -> column tile 2kt (+1)
            const int x0 = v[nt >> 1][2 * kt][2 * (nt & 1)], y0 = v[nt >> 1][2 * kt][2 * (nt & 1) + 1];
            const int x1 = v[nt >> 1][2 * kt + 1][2 * (nt & 1)], y1 = v[nt >> 1][2 * kt + 1][2 * (nt & 1) + 1];
            const uint32_t bH0 = pack_h2(x0 >> 8, y0 >> 8), bL0 = pack_h2(x0 & 255, y0 & 255);
            const uint32_t bH1 = pack_h2(x1 >> 8, y1 >> 8), bL1 = pack_h2(x1 & 255, y1 & 255);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              uint32_t af[4];
#pragma unroll
              for (int r = 0; r < 4; ++r) af[r] = sAVT[((mt * KT + kt) * 4 + r) * 32 + lane];
              mma16816(accH[mt][nt], af, bH0, bH1);
              mma16816(accL[mt][nt], af, bL0, bL1);
            }
          }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int y = __float2int_rn(accH[mt][nt][c]) * 256 + __float2int_rn(accL[mt][nt][c]);
              u[mt][nt][c] = clip3(-32768, 32767, (y + add_i1) >> p.shift_i1);
            }
      }
      // ---- I2: R[iv][ih] = sum_kh U[iv][kh] Th[kh][ih] ---------------------------------------
      {
        float accH[MT][NT][4], accL[MT][NT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int c = 0; c < 4; ++c) { accH[mt][nt][c] = 0.f; accL[mt][nt][c] = 0.f; }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int kt = 0; kt < KT; ++kt) {
            uint32_t aH[4], aL[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int x = u[mt][2 * kt + (r >> 1)][2 * (r & 1)], y = u[mt][2 * kt + (r >> 1)][2 * (r & 1) + 1];
              aH[r] = pack_h2(x >> 8, y >> 8); aL[r] = pack_h2(x & 255, y & 255);
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              const uint32_t b0 = sBH[((kt * NT + nt) * 2 + 0) * 32 + lane], b1 = sBH[((kt * NT + nt) * 2 + 1) * 32 + lane];
              mma16816(accH[mt][nt], aH, b0, b1);
              mma16816(accL[mt][nt], aL, b0, b1);
            }
          }
        // reconstruction: rows mt*16+g (c0,c1) and +8 (c2,c3); segment q: tiles 2q, 2q+1 -> pixels 16q+4t..+3
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh)
#pragma unroll
            for (int qs = 0; qs < KT; ++qs) {
              int px[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int nt = 2 * qs + (e >> 1), c = 2 * hh + (e & 1);
                const int y = __float2int_rn(accH[mt][nt][c]) * 256 + __float2int_rn(accL[mt][nt][c]);
                const int r = clip3(-32768, 32767, (y + add_i2) >> p.shift_i2);
                px[e] = clip3(0, q.pixel_max, (int)(short)(r + predw[2 * mt + hh][qs].at(e)));
                const int df = refw[2 * mt + hh][qs].at(e) - px[e];
                ssd += (unsigned)(df * df);
              }
              Px4<Pix> o;
              o.set(px[0], px[1], px[2], px[3]);
              o.store(prec + (size_t)(mt * 16 + hh * 8 + g) * a.rec.stride + qs * 16 + 4 * t);
            }
      }
    } else {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
          if (!a.rec_is_pred) predw[nt][kt].store(prec + (size_t)(nt * 8 + g) * a.rec.stride + kt * 16 + 4 * t);
#pragma unroll
          for (int e = 0; e < 4; ++e) { const int df = refw[nt][kt].at(e) - predw[nt][kt].at(e); ssd += (unsigned)(df * df); }
        }
    }
    if (a.ssd) {
      ssd = (unsigned)warp_sum((int)ssd);
      if (lane == 0) a.ssd[tu] = (uint32_t)((int32_t)ssd >> q.ssd_shift);
    }
  }
}

template <typename Pix, int N>
static int launch_mma(const FusedArgs &a, size_t n_tus, const QParams &q, const MmaParams &p, cudaStream_t st)
{
  const int warps_per_block = 4;
  size_t grid = (n_tus + warps_per_block - 1) / warps_per_block;
  const size_t cap = (size_t)sm_count() * 8;      // persistent-ish: every CTA builds the fragment tables once
  if (grid > cap) grid = cap;
  fused_mma_kernel<Pix, N><<<(unsigned)grid, warps_per_block * 32, 0, st>>>(a, n_tus, q, p);
  UVGB_LAUNCH_CHECK();
  return 0;
}

int launch_fused_mma(const FusedArgs &a, size_t n_tus, const QParams &q, int bitdepth, int th, int tv, int skip_w,
                     int skip_h, cudaStream_t st)
{
  if (q.w != q.h || (q.w != 16 && q.w != 32) || q.signhide || q.lfnst_idx || a.use_trskip) return -1;
  if (bitdepth > 10) return -1;          // |residual| must stay below 2048 for exact f16 operands
  if (q.q_bits < 15 || q.q_bits > 31) return -1;   // the kernel drops the level clip
  const int N = q.w, l2 = ilog2_host(N);
  MmaParams p;
  p.th = tr_matrix_dev(th, N); p.tv = tr_matrix_dev(tv, N);
  p.shift_f1 = l2 - 1 + bitdepth - 8; p.shift_f2 = l2 + 6; p.shift_i1 = 7; p.shift_i2 = 20 - bitdepth;
  p.cut1 = (th == UVG_B200_DCT2) ? N : N - skip_w;     // 8/16/32-point DST7/DCT8 forward kernels
  p.cut2 = (tv == UVG_B200_DCT2) ? N : N - skip_h;
  if (bitdepth == 8) return N == 16 ? launch_mma<uint8_t, 16>(a, n_tus, q, p, st) : launch_mma<uint8_t, 32>(a, n_tus, q, p, st);
  return N == 16 ? launch_mma<uint16_t, 16>(a, n_tus, q, p, st) : launch_mma<uint16_t, 32>(a, n_tus, q, p, st);
}

}  // namespace uvgb
