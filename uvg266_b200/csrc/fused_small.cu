// fused_small.cu - quantize_residual for square 4x4 and 8x8 TUs: ONE THREAD PER TU, the whole
// chain (residual -> forward transform -> quant -> dequant -> inverse transform -> recon ->
// SSD / coefficient cost) stays in registers: no shared memory, no barriers.
//
// Why: 4x4 and 8x8 TUs are 92% of the TUs of an all-intra CTU and a
// 16/64-sample block is too small to share between threads without spending more on
// synchronisation than on arithmetic. With dense operands thread t touches bytes
// [t*N*N, (t+1)*N*N) of each array, so a warp's loads and stores are fully coalesced.
//
// Arithmetic restates quant-generic.c:460-612 with the matrix form of the transforms
// (transform.cuh); the transform matrices arrive as kernel parameters (constant bank), so
// any DCT-2/DST-7/DCT-8 pair runs through the same code.
#include "quant.cuh"

namespace uvgb {

template <int N>
struct SmallTr {
  int th[N * N];     // horizontal matrix T[k*N+i]
  int tv[N * N];     // vertical matrix
  int shift_f1, shift_f2, shift_i1, shift_i2;
};

// A block lives in registers as packed words (row-major pixels, 4 or 2 per word): a quarter /
// half of the registers an int-per-pixel array would need, which is what lets an 8x8 TU stay
// below the 255-register limit.
template <typename Pix, int N> struct BlockWords {
  static constexpr int PW = 4 / (int)sizeof(Pix);
  static constexpr int NWORDS = N * N / PW;
  uint32_t w[NWORDS];
  __device__ __forceinline__ int at(int i) const {
    return (int)((w[i / PW] >> ((i % PW) * (32 / PW))) & ((1u << (32 / PW)) - 1u));
  }
};

template <typename Pix, int N>
__device__ __forceinline__ void load_block(const Pix *p, uint32_t stride, BlockWords<Pix, N> &b)
{
  constexpr int PW = BlockWords<Pix, N>::PW;
  constexpr int NW = N / PW;                       // words per row
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  if (stride == (uint32_t)N && (a & 15) == 0 && (N * N * sizeof(Pix)) % 16 == 0) {
    constexpr int NV = (N * N * (int)sizeof(Pix)) / 16;   // dense, 16-byte aligned: vector loads
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p) + i);
      b.w[4 * i] = q.x; b.w[4 * i + 1] = q.y; b.w[4 * i + 2] = q.z; b.w[4 * i + 3] = q.w;
    }
    return;
  }
#pragma unroll
  for (int y = 0; y < N; ++y) {
    const Pix *r = p + (size_t)y * stride;
    if ((reinterpret_cast<uintptr_t>(r) & 3) == 0) {
#pragma unroll
      for (int i = 0; i < NW; ++i) b.w[y * NW + i] = __ldg(reinterpret_cast<const uint32_t *>(r) + i);
    } else {
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < PW; ++j) v |= (uint32_t)__ldg(r + i * PW + j) << (j * (32 / PW));
        b.w[y * NW + i] = v;
      }
    }
  }
}

template <typename Pix, int N>
__device__ __forceinline__ void store_block(Pix *p, uint32_t stride, const BlockWords<Pix, N> &b)
{
  constexpr int PW = BlockWords<Pix, N>::PW;
  constexpr int NW = N / PW;
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  if (stride == (uint32_t)N && (a & 15) == 0 && (N * N * sizeof(Pix)) % 16 == 0) {
    constexpr int NV = (N * N * (int)sizeof(Pix)) / 16;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      reinterpret_cast<uint4 *>(p)[i] = make_uint4(b.w[4 * i], b.w[4 * i + 1], b.w[4 * i + 2], b.w[4 * i + 3]);
    return;
  }
#pragma unroll
  for (int y = 0; y < N; ++y) {
    Pix *r = p + (size_t)y * stride;
    if ((reinterpret_cast<uintptr_t>(r) & 3) == 0) {
#pragma unroll
      for (int i = 0; i < NW; ++i) reinterpret_cast<uint32_t *>(r)[i] = b.w[y * NW + i];
    } else {
#pragma unroll
      for (int x = 0; x < N; ++x) r[x] = (Pix)b.at(y * N + x);
    }
  }
}

template <typename Pix, int N>
__global__ void __launch_bounds__(128)
fused_small_kernel(FusedArgs a, size_t n_tus, QParams q, SmallTr<N> tr)
{
  const size_t tu = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tu >= n_tus) return;
  const Pix *pref = static_cast<const Pix *>(a.ref.base) + a.ref.offset(tu);
  const Pix *ppred = static_cast<const Pix *>(a.pred.base) + a.pred.offset(tu);
  Pix *prec = const_cast<Pix *>(static_cast<const Pix *>(a.rec.base)) + a.rec.offset(tu);

  BlockWords<Pix, N> refb, predb;
  load_block<Pix, N>(pref, a.ref.stride, refb);
  load_block<Pix, N>(ppred, a.pred.stride, predb);

  // residual (generate_residual) and forward transform: x -> t (pass 1, transposed) -> c
  int c[N * N];
  {
    int t[N * N];
    const int add1 = tr.shift_f1 > 0 ? 1 << (tr.shift_f1 - 1) : 0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      int x[N];
#pragma unroll
      for (int i = 0; i < N; ++i) x[i] = refb.at(j * N + i) - predb.at(j * N + i);
#pragma unroll
      for (int k = 0; k < N; ++k) {
        int s = add1;
#pragma unroll
        for (int i = 0; i < N; ++i) s += x[i] * tr.th[k * N + i];
        t[k * N + j] = (int)(short)(s >> tr.shift_f1);       // dst[k*line + j], int16 wrap
      }
    }
    if (a.use_trskip) {
#pragma unroll
      for (int i = 0; i < N * N; ++i) c[i] = refb.at(i) - predb.at(i);   // uvg_transformskip: coeff = residual
    } else {
      const int add2 = 1 << (tr.shift_f2 - 1);
#pragma unroll
      for (int j = 0; j < N; ++j)          // j = horizontal frequency (row of t)
#pragma unroll
        for (int k = 0; k < N; ++k) {      // k = vertical frequency
          int s = add2;
#pragma unroll
          for (int i = 0; i < N; ++i) s += t[j * N + i] * tr.tv[k * N + i];
          c[k * N + j] = (int)(short)(s >> tr.shift_f2);
        }
    }
  }

  // quant (uvg_quant_generic level loop, no sign hiding on this path), coefficient cost
  int qc[N * N];
  unsigned nz = 0, cost = 0;
  const uint32_t wlo = a.w0 | (a.w1 << 16), whi = a.w2 | (a.w3 << 16);
#pragma unroll
  for (int i = 0; i < N * N; ++i) {
    // no int16 clip of the level: the launcher admits q_bits >= 15 only, so |level| <= (32768 * 26214 + add) >> 15 < 32767
    const unsigned ul = ((unsigned)abs(c[i]) * (unsigned)q.coeff_default + (unsigned)q.add) >> q.q_bits;
    qc[i] = c[i] < 0 ? -(int)ul : (int)ul;
    nz |= ul;
    // weight of min(|level|, 3) from the packed Q8.8 table; bytes 2, 3 of the PRMT result are copies of the table's byte
    // 0 - a constant, taken off once after the loop
    cost += __byte_perm(wlo, whi, min(ul, 3u) * 0x22u + 0x10u);
  }
  cost -= ((wlo & 0xffu) * 0x01010000u) * (unsigned)(N * N);
  // coefficients out: N*N int16 contiguous per TU
  {
    uint32_t *dst = reinterpret_cast<uint32_t *>(a.coeff_out + tu * (size_t)(N * N));
#pragma unroll
    for (int i = 0; i < N * N / 8; ++i) {
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        w[k] = __byte_perm((uint32_t)qc[8 * i + 2 * k], (uint32_t)qc[8 * i + 2 * k + 1], 0x5410);
      reinterpret_cast<uint4 *>(dst)[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }

  const bool recon = nz && !a.early_skip;
  BlockWords<Pix, N> recb = predb;
  if (recon) {
    // dequant (flat) -> inverse transform -> reconstruction
    int r[N * N];
    if (a.use_trskip) {
#pragma unroll
      for (int i = 0; i < N * N; ++i)
        r[i] = clip3(-32768, 32767, (qc[i] * q.dq_scale + q.dq_add) >> q.dq_shift);
    } else {
      int u[N * N];
      const int add1 = 1 << (tr.shift_i1 - 1);
#pragma unroll
      for (int j = 0; j < N; ++j) {        // j = horizontal frequency (column of the coefficients)
        int d[N];
#pragma unroll
        for (int k = 0; k < N; ++k) d[k] = clip3(-32768, 32767, (qc[k * N + j] * q.dq_scale + q.dq_add) >> q.dq_shift);
#pragma unroll
        for (int i = 0; i < N; ++i) {      // i = vertical sample
          int s = add1;
#pragma unroll
          for (int k = 0; k < N; ++k) s += d[k] * tr.tv[k * N + i];
          u[j * N + i] = clip3(-32768, 32767, s >> tr.shift_i1);
        }
      }
      const int add2 = 1 << (tr.shift_i2 - 1);
#pragma unroll
      for (int j = 0; j < N; ++j)          // j = vertical sample
#pragma unroll
        for (int i = 0; i < N; ++i) {      // i = horizontal sample
          int s = add2;
#pragma unroll
          for (int k = 0; k < N; ++k) s += u[k * N + j] * tr.th[k * N + i];
          r[j * N + i] = clip3(-32768, 32767, s >> tr.shift_i2);
        }
    }
    constexpr int PW = BlockWords<Pix, N>::PW;
#pragma unroll
    for (int wi = 0; wi < BlockWords<Pix, N>::NWORDS; ++wi) {
      if constexpr (sizeof(Pix) == 1) {                      // 8-bit: clip to 0..255 and pack in two saturating packs
        int t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = (int)(short)(r[wi * 4 + j] + predb.at(wi * 4 + j));
        recb.w[wi] = pack4_sat_u8(t[0], t[1], t[2], t[3]);
      } else {
        uint32_t x = 0;
#pragma unroll
        for (int j = 0; j < PW; ++j) {
          const int i = wi * PW + j;
          x |= (uint32_t)clip3(0, q.pixel_max, (int)(short)(r[i] + predb.at(i))) << (j * (32 / PW));
        }
        recb.w[wi] = x;
      }
    }
  }
  if (recon || !a.rec_is_pred) store_block<Pix, N>(prec, a.rec.stride, recb);
  a.has_coeffs[tu] = nz ? 1 : 0;
  if (a.ssd) {
    unsigned ssd = 0;
    if constexpr (sizeof(Pix) == 1) {
#pragma unroll
      for (int wi = 0; wi < BlockWords<Pix, N>::NWORDS; ++wi) { const uint32_t d4 = __vabsdiffu4(refb.w[wi], recb.w[wi]); ssd = __dp4a(d4, d4, ssd); }
    } else {
#pragma unroll
      for (int i = 0; i < N * N; ++i) { const int df = refb.at(i) - recb.at(i); ssd += (unsigned)(df * df); }
    }
    a.ssd[tu] = (uint32_t)((int32_t)ssd >> q.ssd_shift);
  }
  if (a.cost) a.cost[tu] = (cost + 128u) >> 8;
}

template <typename Pix, int N>
static int launch_small(const FusedArgs &a, size_t n_tus, const QParams &q, int bitdepth, int th, int tv,
                        cudaStream_t st)
{
  SmallTr<N> tr;
  const int8_t *mh = tr_matrix_host(th, N), *mv = tr_matrix_host(tv, N);
  for (int i = 0; i < N * N; ++i) { tr.th[i] = mh[i]; tr.tv[i] = mv[i]; }
  const int l2 = ilog2_host(N);
  tr.shift_f1 = l2 - 1 + bitdepth - 8; tr.shift_f2 = l2 + 6;
  tr.shift_i1 = 7; tr.shift_i2 = 20 - bitdepth;
  const int block = 128;
  const unsigned grid = (unsigned)((n_tus + block - 1) / block);
  fused_small_kernel<Pix, N><<<grid, block, 0, st>>>(a, n_tus, q, tr);
  UVGB_LAUNCH_CHECK();
  return 0;
}

int launch_fused_small(const FusedArgs &a, size_t n_tus, const QParams &q, int bitdepth, int th, int tv,
                       cudaStream_t st)
{
  if (q.w != q.h || (q.w != 4 && q.w != 8) || q.signhide || q.lfnst_idx) return -1;
  if (q.q_bits < 15 || q.q_bits > 31) return -1;           // the kernel drops the level clip (see the quant loop)
  if (bitdepth == 8) {
    return q.w == 4 ? launch_small<uint8_t, 4>(a, n_tus, q, bitdepth, th, tv, st)
                    : launch_small<uint8_t, 8>(a, n_tus, q, bitdepth, th, tv, st);
  }
  return q.w == 4 ? launch_small<uint16_t, 4>(a, n_tus, q, bitdepth, th, tv, st)
                  : launch_small<uint16_t, 8>(a, n_tus, q, bitdepth, th, tv, st);
}

}  // namespace uvgb
