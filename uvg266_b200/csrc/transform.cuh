// transform.cuh - device building blocks shared by the standalone transform kernels
// (transform.cu) and the fused quantize_residual kernel (quant.cu).
//
// Matrix formulation of uvg266's separable integer transforms (SURVEY.md Appendix D,
// verified bit-exact against mts_dct_generic / mts_idct_generic and the partial-butterfly
// dct_NxN / idct_NxN of dct-generic.c): int32 accumulation, per-pass rounding shift,
// forward results converted to int16 by plain C conversion (wrap), inverse results
// clipped to int16.
#pragma once
#include "common.cuh"

namespace uvgb {

// Everything a 2-D transform of one shape needs; filled on the host.
struct TrPlan {
  int w, h;
  int n1, line1, reduced1, cutoff1, shift1;   // first pass
  int n2, line2, reduced2, cutoff2, shift2;   // second pass (n2 == 0: single-pass 1-D case)
  const int8_t *t1;                           // device matrices, row-major T[k*n+i]
  const int8_t *t2;
  int th, tv;                                 // transform types (UVG_B200_DCT2 / DST7 / DCT8) of the two directions
};

// Forward kernel contract of the reference, e.g. fastForwardDCT2_B8 (dct-generic.c:1159),
// fastForwardDST7_B16 (:1574): transform `reduced` rows of n samples, write transposed
// (dst[k*line + j]); rows >= reduced and frequencies >= cutoff are zero.
// `T` is in shared memory, transposed (Tt[i*n + k]) so that lanes with consecutive k read
// consecutive addresses.
__device__ __forceinline__ void fwd_pass(const int16_t *__restrict__ src, int16_t *__restrict__ dst,
                                         const int8_t *__restrict__ Tt, int n, int line,
                                         int reduced, int cutoff, int shift, int tid, int nthreads)
{
  const int add = shift > 0 ? 1 << (shift - 1) : 0;
  if (n >= 16) {             // (n % 4 == 0 and enough outputs per TU to keep every thread busy)
    // four consecutive outputs per thread: one 32-bit word of the transposed matrix (4 int8) and one
    // data sample per contraction step feed 4 MACs (the matrix rows are 4-byte aligned for n % 4 == 0)
    const int nq = n >> 2, total4 = nq * line;
    for (int o = tid; o < total4; o += nthreads) {
      const int j = o / nq, k0 = (o - j * nq) << 2;
      int sum[4] = { 0, 0, 0, 0 };
      if (j < reduced && k0 < cutoff) {
        const int16_t *s = src + j * n;
#pragma unroll 4
        for (int i = 0; i < n; ++i) {
          const int sv = s[i];
          const uint32_t w = *reinterpret_cast<const uint32_t *>(Tt + i * n + k0);
          sum[0] += sv * (int)(signed char)(w & 0xff); sum[1] += sv * (int)(signed char)((w >> 8) & 0xff);
          sum[2] += sv * (int)(signed char)((w >> 16) & 0xff); sum[3] += sv * ((int)w >> 24);
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = k0 + c;
        dst[k * line + j] = (j < reduced && k < cutoff) ? (int16_t)(short)((sum[c] + add) >> shift) : (int16_t)0;
      }
    }
    return;
  }
  const int total = n * line;
  for (int o = tid; o < total; o += nthreads) {
    const int j = o / n, k = o - j * n;      // k fastest: conflict-free reads of Tt
    int v = 0;
    if (j < reduced && k < cutoff) {
      int sum = 0;
      const int16_t *s = src + j * n;
#pragma unroll 4
      for (int i = 0; i < n; ++i) sum += (int)s[i] * (int)Tt[i * n + k];
      v = (int)(short)((sum + add) >> shift);
    }
    dst[k * line + j] = (int16_t)v;
  }
}

// Inverse kernel contract (fastInverseDCT2_B8 :1207 ...): dst[j*n + i] =
// clip16(rnd(sum_{k<cutoff} src[k*line + j] * T[k*n + i])); rows >= reduced are zero.
// `T` in shared memory row-major (T[k*n + i]): lanes with consecutive i read consecutive
// addresses.
__device__ __forceinline__ void inv_pass(const int16_t *__restrict__ src, int16_t *__restrict__ dst,
                                         const int8_t *__restrict__ T, int n, int line,
                                         int reduced, int cutoff, int shift, int tid, int nthreads)
{
  const int add = shift > 0 ? 1 << (shift - 1) : 0;
  if (n >= 16) {             // (n % 4 == 0 and enough outputs per TU to keep every thread busy)
    const int nq = n >> 2, total4 = nq * line;
    for (int o = tid; o < total4; o += nthreads) {
      const int j = o / nq, i0 = (o - j * nq) << 2;
      int sum[4] = { 0, 0, 0, 0 };
      if (j < reduced) {
#pragma unroll 4
        for (int k = 0; k < cutoff; ++k) {
          const int sv = src[k * line + j];
          const uint32_t w = *reinterpret_cast<const uint32_t *>(T + k * n + i0);
          sum[0] += sv * (int)(signed char)(w & 0xff); sum[1] += sv * (int)(signed char)((w >> 8) & 0xff);
          sum[2] += sv * (int)(signed char)((w >> 16) & 0xff); sum[3] += sv * ((int)w >> 24);
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        dst[j * n + i0 + c] = j < reduced ? (int16_t)clip3(-32768, 32767, (sum[c] + add) >> shift) : (int16_t)0;
    }
    return;
  }
  const int total = n * line;
  for (int o = tid; o < total; o += nthreads) {
    const int j = o / n, i = o - j * n;
    int v = 0;
    if (j < reduced) {
      int sum = 0;
#pragma unroll 4
      for (int k = 0; k < cutoff; ++k) sum += (int)src[k * line + j] * (int)T[k * n + i];
      v = clip3(-32768, 32767, (sum + add) >> shift);
    }
    dst[j * n + i] = (int16_t)v;
  }
}

// Host-side plan construction (transform.cu).
int make_fwd_plan(TrPlan &p, int bitdepth, int w, int h, int type_hor, int type_ver, int skip_w,
                  int skip_h);
int make_inv_plan(TrPlan &p, int bitdepth, int w, int h, int type_hor, int type_ver, int skip_w,
                  int skip_h);
// square 16x16 / 32x32 on the INT8 tensor cores (transform_imma.cu). 0: launched; -1: not covered; 1: error
int launch_transform_imma(bool inverse, const TrPlan &p, const int16_t *in, int16_t *out, size_t n_tus, cudaStream_t st);
// uvg_get_tr_type + the zero-out rules of mts_dct_generic (dct-generic.c:2501-2600).
void derive_tr(int w, int h, int color, int cu_type, int tr_idx, int lfnst_idx, int cr_lfnst_idx,
               int isp_mode, int mts_type, int *type_hor, int *type_ver, int *skip_w, int *skip_h);

}  // namespace uvgb
