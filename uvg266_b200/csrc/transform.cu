// transform.cu - DCT-2 / DST-7 / DCT-8 (MTS) forward and inverse transforms, batched over
// TUs, for sm_100a. Restates dct-generic.c (dct_NxN :720, idct_NxN :731, mts_dct_generic
// :2560, mts_idct_generic :2621, uvg_get_tr_type :2501) in matrix form.
#include "transform.cuh"
#include <mutex>
#include <string.h>

namespace uvgb {

// -------------------------------------------------------------------------------------
// Transform matrices from the VVC base constants
// -------------------------------------------------------------------------------------
// DCT-2: tuned round(64*sqrt(2)*cos(m*pi/64)), m = 0..31 (m = 0 stands for the 64 of the
// DC row) - the values of the reference's uvg_g_DCT2P32 (dct-generic.c:1011-1015).
static const int8_t k_dct2_mag[32] = { 64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73,
  70, 67, 64, 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4 };
// DST-7: c_j ~ sin(j*pi/(2N+1)), j = 1..N, as tuned by VVC (uvg_g_DST7P* :1018-1022).
static const int8_t k_dst7_mag4[4] = { 29, 55, 74, 84 };
static const int8_t k_dst7_mag8[8] = { 17, 32, 46, 60, 71, 78, 85, 86 };
static const int8_t k_dst7_mag16[16] = { 8, 17, 25, 33, 40, 48, 55, 62, 68, 73, 77, 81, 85, 87, 88, 88 };
static const int8_t k_dst7_mag32[32] = { 4, 9, 13, 17, 21, 26, 30, 34, 38, 42, 46, 50, 53, 56,
  60, 63, 66, 68, 72, 74, 77, 78, 80, 82, 84, 85, 86, 87, 88, 89, 90, 90 };

static int dct2_entry(int n, int k, int i)
{
  if (k == 0) return 64;
  int m = ((2 * i + 1) * k * (32 / n)) % 128;   // cos(m*pi/64)
  if (m > 64) m = 128 - m;
  if (m == 32) return 0;
  return m < 32 ? k_dct2_mag[m] : -k_dct2_mag[64 - m];
}

static int dst7_entry(int n, int k, int i)
{
  const int8_t *mag = n == 4 ? k_dst7_mag4 : n == 8 ? k_dst7_mag8 : n == 16 ? k_dst7_mag16 : k_dst7_mag32;
  const int p = 2 * n + 1;
  int m = ((2 * k + 1) * (i + 1)) % (2 * p);    // sin(m*pi/p)
  int sign = 1;
  if (m > p) { m -= p; sign = -1; }
  if (m > n) m = p - m;
  return m == 0 ? 0 : sign * mag[m - 1];
}

// [type][log2n] -> offset into one int8 pool; log2n = 1..5
static int8_t h_pool[3 * (4 + 16 + 64 + 256 + 1024)];
static int h_off[3][6];
static int8_t *d_pool = nullptr;
static std::once_flag tables_once;
static cudaError_t tables_err = cudaSuccess;

static void build_tables()
{
  int off = 0;
  for (int t = 0; t < 3; ++t)
    for (int l = 1; l <= 5; ++l) {
      const int n = 1 << l;
      h_off[t][l] = off;
      for (int k = 0; k < n; ++k)
        for (int i = 0; i < n; ++i) {
          int v;
          if (t == UVG_B200_DCT2) v = dct2_entry(n, k, i);
          else if (n == 2) v = dct2_entry(n, k, i);             // no 2-point DST7/DCT8 exists
          else if (t == UVG_B200_DST7) v = dst7_entry(n, k, i);
          else v = ((k & 1) ? -1 : 1) * dst7_entry(n, k, n - 1 - i);  // DCT-8 = mirrored DST-7
          h_pool[off + k * n + i] = (int8_t)v;
        }
      off += n * n;
    }
  tables_err = cudaMalloc(&d_pool, sizeof(h_pool));
  if (tables_err == cudaSuccess)
    tables_err = cudaMemcpy(d_pool, h_pool, sizeof(h_pool), cudaMemcpyHostToDevice);
}

int init_transform_tables()
{
  std::call_once(tables_once, build_tables);
  if (tables_err != cudaSuccess) { set_error("transform tables: %s", cudaGetErrorString(tables_err)); return 1; }
  return 0;
}

const int8_t *tr_matrix_dev(int type, int n) { return d_pool + h_off[type][ilog2_host(n)]; }
const int8_t *tr_matrix_host(int type, int n) { return h_pool + h_off[type][ilog2_host(n)]; }

// -------------------------------------------------------------------------------------
// plans
// -------------------------------------------------------------------------------------
static bool valid_dim(int d) { return d == 1 || d == 2 || d == 4 || d == 8 || d == 16 || d == 32; }

static int check_tr(int w, int h, int th, int tv)
{
  if (!valid_dim(w) || !valid_dim(h) || (w == 1 && h == 1)) { set_error("transform %dx%d unsupported", w, h); return 1; }
  if (th < 0 || th > 2 || tv < 0 || tv > 2) { set_error("bad transform type"); return 1; }
  // dct_table / idct_table have no 2-point DST7/DCT8 (dct-generic.c:2484-2494)
  if ((w < 4 && th != UVG_B200_DCT2) || (h < 4 && tv != UVG_B200_DCT2)) { set_error("no DST7/DCT8 below 4 points"); return 1; }
  return 0;
}

int make_fwd_plan(TrPlan &p, int bd, int w, int h, int th, int tv, int skip_w, int skip_h)
{
  if (check_tr(w, h, th, tv)) return 1;
  const int l2w = ilog2_host(w), l2h = ilog2_host(h);
  // only the 8/16/32-point DST7/DCT8 forward kernels honour skip_line2 (:1574-2343)
  const int cut_w = (th == UVG_B200_DCT2 || w < 8) ? w : w - skip_w;
  const int cut_h = (tv == UVG_B200_DCT2 || h < 8) ? h : h - skip_h;
  p.w = w; p.h = h; p.th = th; p.tv = tv;
  if (h == 1) {            // :2608
    p.n1 = w; p.line1 = 1; p.reduced1 = 1; p.cutoff1 = cut_w; p.shift1 = l2w - 1 + bd - 8;
    p.t1 = tr_matrix_dev(th, w); p.n2 = 0; p.t2 = nullptr;
  } else if (w == 1) {     // :2610-2611, shift taken from VTM
    p.n1 = h; p.line1 = 1; p.reduced1 = 1; p.cutoff1 = cut_h; p.shift1 = l2h + bd - 9;
    p.t1 = tr_matrix_dev(tv, h); p.n2 = 0; p.t2 = nullptr;
  } else {                 // :2613-2615
    p.n1 = w; p.line1 = h; p.reduced1 = h; p.cutoff1 = cut_w; p.shift1 = l2w - 1 + bd - 8;
    p.t1 = tr_matrix_dev(th, w);
    p.n2 = h; p.line2 = w; p.reduced2 = w - skip_w; p.cutoff2 = cut_h; p.shift2 = l2h + 6;
    p.t2 = tr_matrix_dev(tv, h);
  }
  return 0;
}

int make_inv_plan(TrPlan &p, int bd, int w, int h, int th, int tv, int skip_w, int skip_h)
{
  if (check_tr(w, h, th, tv)) return 1;
  // only the 8-point DST7/DCT8 inverse kernels stop at N - skip_line2 (:2272, :2343)
  const int cut_w = (th != UVG_B200_DCT2 && w == 8) ? w - skip_w : w;
  const int cut_h = (tv != UVG_B200_DCT2 && h == 8) ? h - skip_h : h;
  const int shift2 = 20 - bd;
  p.w = w; p.h = h; p.th = th; p.tv = tv;
  if (h == 1) {            // :2669
    p.n1 = w; p.line1 = 1; p.reduced1 = 1; p.cutoff1 = cut_w; p.shift1 = shift2 + 1;
    p.t1 = tr_matrix_dev(th, w); p.n2 = 0; p.t2 = nullptr;
  } else if (w == 1) {     // :2671
    p.n1 = h; p.line1 = 1; p.reduced1 = 1; p.cutoff1 = cut_h; p.shift1 = shift2 + 1;
    p.t1 = tr_matrix_dev(tv, h); p.n2 = 0; p.t2 = nullptr;
  } else {                 // :2673-2674: vertical first, then horizontal
    p.n1 = h; p.line1 = w; p.reduced1 = w - skip_w; p.cutoff1 = cut_h; p.shift1 = 7;
    p.t1 = tr_matrix_dev(tv, h);
    p.n2 = w; p.line2 = h; p.reduced2 = h; p.cutoff2 = cut_w; p.shift2 = shift2;
    p.t2 = tr_matrix_dev(th, w);
  }
  return 0;
}

// uvg_get_tr_type (dct-generic.c:2501-2557) + skip derivation of mts_dct_generic
// (:2582-2600). enum values: cu_type CU_INTRA=1/CU_INTER=2 (cu.h:50), mts_type = enum
// uvg_mts (uvg266.h:224), tr_idx = mts_idx (cu.h:70).
void derive_tr(int w, int h, int color, int cu_type, int tr_idx, int lfnst_idx, int cr_lfnst_idx,
               int isp_mode, int mts_type, int *th, int *tv, int *skip_w, int *skip_h)
{
  *th = UVG_B200_DCT2; *tv = UVG_B200_DCT2;
  if (color == 0) {
    const bool intra = cu_type == 1;
    const bool explicit_mts = mts_type == 3 || (intra ? mts_type == 1 : (mts_type == 2 && cu_type == 2));
    const bool implicit_mts = intra && (mts_type == 4 || mts_type == 2);
    const bool is_isp = intra && isp_mode;
    if (!(is_isp && lfnst_idx)) {
      if (implicit_mts || (is_isp && explicit_mts)) {
        if (w >= 4 && w <= 16) *th = UVG_B200_DST7;
        if (h >= 4 && h <= 16) *tv = UVG_B200_DST7;
      } else if (explicit_mts && tr_idx > 1) {
        static const int subset[4][2] = { { UVG_B200_DST7, UVG_B200_DST7 }, { UVG_B200_DCT8, UVG_B200_DST7 },
                                          { UVG_B200_DST7, UVG_B200_DCT8 }, { UVG_B200_DCT8, UVG_B200_DCT8 } };
        *th = subset[tr_idx - 2][0]; *tv = subset[tr_idx - 2][1];
      }
    }
  }
  *skip_w = (*th != UVG_B200_DCT2 && w == 32) ? 16 : (w > 32 ? w - 32 : 0);
  *skip_h = (*tv != UVG_B200_DCT2 && h == 32) ? 16 : (h > 32 ? h - 32 : 0);
  if ((lfnst_idx && color == 0) || (cr_lfnst_idx && color != 0)) {
    if ((w == 4 && h > 4) || (w > 4 && h == 4)) { *skip_w = w - 4; *skip_h = h - 4; }
    else if (w >= 8 && h >= 8) { *skip_w = w - 8; *skip_h = h - 8; }
  }
}

// -------------------------------------------------------------------------------------
// standalone batched kernel
// -------------------------------------------------------------------------------------
// A CTA processes SLOTS TUs at a time (SLOTS * w*h <= 1024 outputs per pass, 256 threads).
// Shared memory: both matrices (forward: transposed; inverse: row-major) + per-slot src
// and tmp buffers.
template <bool INV>
__global__ void __launch_bounds__(256)
transform_kernel(const int16_t *__restrict__ in, int16_t *__restrict__ out, size_t n_tus, TrPlan p,
                 int slots)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int area = p.w * p.h;
  int8_t *sT1 = reinterpret_cast<int8_t *>(smem_raw);
  int8_t *sT2 = sT1 + p.n1 * p.n1;
  int16_t *sbuf = reinterpret_cast<int16_t *>(smem_raw + ((p.n1 * p.n1 + p.n2 * p.n2 + 15) & ~15));
  // matrices
  for (int t = threadIdx.x; t < p.n1 * p.n1; t += blockDim.x) {
    const int k = t / p.n1, i = t - k * p.n1;
    if (INV) sT1[t] = p.t1[t]; else sT1[i * p.n1 + k] = p.t1[t];
  }
  for (int t = threadIdx.x; t < p.n2 * p.n2; t += blockDim.x) {
    const int k = t / p.n2, i = t - k * p.n2;
    if (INV) sT2[t] = p.t2[t]; else sT2[i * p.n2 + k] = p.t2[t];
  }
  const int tps = blockDim.x / slots;             // threads per slot
  const int slot = threadIdx.x / tps, tid = threadIdx.x - slot * tps;
  int16_t *s_src = sbuf + (size_t)slot * 2 * area;
  int16_t *s_tmp = s_src + area;
  for (size_t base = (size_t)blockIdx.x * slots; base < n_tus; base += (size_t)gridDim.x * slots) {
    const size_t tu = base + slot;
    const bool active = tu < n_tus;
    __syncthreads();                              // matrices ready / previous round consumed
    if (active)
      for (int t = tid; t < area; t += tps) s_src[t] = in[tu * area + t];
    __syncthreads();
    if (active) {
      if (INV) inv_pass(s_src, p.n2 ? s_tmp : s_src + area, sT1, p.n1, p.line1, p.reduced1, p.cutoff1, p.shift1, tid, tps);
      else fwd_pass(s_src, s_tmp, sT1, p.n1, p.line1, p.reduced1, p.cutoff1, p.shift1, tid, tps);
    }
    __syncthreads();
    if (active) {
      if (p.n2) {
        if (INV) inv_pass(s_tmp, s_src, sT2, p.n2, p.line2, p.reduced2, p.cutoff2, p.shift2, tid, tps);
        else fwd_pass(s_tmp, s_src, sT2, p.n2, p.line2, p.reduced2, p.cutoff2, p.shift2, tid, tps);
      }
    }
    __syncthreads();
    if (active) {
      const int16_t *res = p.n2 ? s_src : s_tmp;
      for (int t = tid; t < area; t += tps) out[tu * area + t] = res[t];
    }
  }
}

// ---- fast 2-D path: every dimension a multiple of 4 ---------------------------------------------------
// Both passes read their contraction index contiguously: data rows as pairs of int16 (one 32-bit word),
// matrix rows as 4 int8 per word, so one IDP.2A (dp2a) does two MACs; a thread owns a 4 x 4 tile of
// outputs (4 matrix words + 4 x 2 data words feed 32 dp2a per step of 4 contraction indices).
// Forward: load natural, pass 1 and pass 2 both write transposed (fastForward* contract).
// Inverse: the coefficient block is transposed on the way into shared memory, pass 1 writes
// transposed, pass 2 natural; the matrices are stored transposed (M[i][k] = T[k][i]).
struct FastPass { int n, contr, out_limit, line, reduced, shift, clip, transposed_out; };

__device__ __forceinline__ void fast_pass(const int16_t *__restrict__ src, int16_t *__restrict__ dst,
                                          const int8_t *__restrict__ M, const FastPass f, int tid, int nthreads)
{
  const int add = f.shift > 0 ? 1 << (f.shift - 1) : 0;
  const int nq = f.n >> 2, lq = f.line >> 2;
  // rows are padded by 4 elements (data) / 4 bytes (matrix): tiles 4 rows apart land in different banks
  const int sp = f.n + 4, mp = f.n + 4, dp = f.transposed_out ? f.line + 4 : f.n + 4;
  for (int o = tid; o < nq * lq; o += nthreads) {
    const int jq = o / nq, k0 = (o - jq * nq) << 2, j0 = jq << 2;
    int acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0;
    if (j0 < f.reduced && k0 < f.out_limit) {
      for (int i = 0; i < f.contr; i += 4) {
        uint32_t m[4];
        uint2 d[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) m[a] = *reinterpret_cast<const uint32_t *>(M + (k0 + a) * mp + i);
#pragma unroll
        for (int b = 0; b < 4; ++b) d[b] = *reinterpret_cast<const uint2 *>(src + (j0 + b) * sp + i);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = __dp2a_hi((int)d[b].y, (int)m[a], __dp2a_lo((int)d[b].x, (int)m[a], acc[a][b]));
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int k = k0 + a, j = j0 + b;
        int v = 0;
        if (j < f.reduced && k < f.out_limit) {
          v = (acc[a][b] + add) >> f.shift;
          v = f.clip ? clip3(-32768, 32767, v) : (int)(short)v;
        }
        if (f.transposed_out) dst[k * dp + j] = (int16_t)v; else dst[j * dp + k] = (int16_t)v;
      }
  }
}

template <bool INV>
__global__ void __launch_bounds__(256)
transform_fast_kernel(const int16_t *__restrict__ in, int16_t *__restrict__ out, size_t n_tus, TrPlan p, int slots,
                      FastPass f1, FastPass f2)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int area = p.w * p.h;
  const int mx = p.w > p.h ? p.w : p.h, bsz = (mx + 4) * mx;          // one padded buffer (int16 elements)
  int8_t *sM1 = reinterpret_cast<int8_t *>(smem_raw);
  int8_t *sM2 = sM1 + p.n1 * (p.n1 + 4);
  int16_t *sbuf = reinterpret_cast<int16_t *>(smem_raw + ((p.n1 * (p.n1 + 4) + p.n2 * (p.n2 + 4) + 15) & ~15));
  for (int t = threadIdx.x; t < p.n1 * p.n1; t += blockDim.x) {
    const int k = t / p.n1, i = t - k * p.n1;
    if (INV) sM1[i * (p.n1 + 4) + k] = p.t1[t]; else sM1[k * (p.n1 + 4) + i] = p.t1[t];
  }
  for (int t = threadIdx.x; t < p.n2 * p.n2; t += blockDim.x) {
    const int k = t / p.n2, i = t - k * p.n2;
    if (INV) sM2[i * (p.n2 + 4) + k] = p.t2[t]; else sM2[k * (p.n2 + 4) + i] = p.t2[t];
  }
  const int tps = blockDim.x / slots, slot = threadIdx.x / tps, tid = threadIdx.x - slot * tps;
  int16_t *bufA = sbuf + (size_t)slot * 2 * bsz, *bufB = bufA + bsz;
  const int ip = f1.n + 4;                          // pitch of the pass-1 input rows
  const int op = (INV ? f2.n : f2.line) + 4;        // pitch of the final rows (forward: transposed write)
  const int ow = INV ? f2.n : f2.line;              // = p.w: both directions end row-major [h][w]
  for (size_t base = (size_t)blockIdx.x * slots; base < n_tus; base += (size_t)gridDim.x * slots) {
    const size_t tu = base + slot;
    const bool active = tu < n_tus;
    __syncthreads();
    if (active) {
      if (INV) {                                   // in[k][j] -> bufA[j][k]: the contraction index becomes contiguous
        for (int t = tid; t < area; t += tps) { const int k = t / f1.line, j = t - k * f1.line; bufA[j * ip + k] = in[tu * area + t]; }
      } else {
        const int vpr = f1.n >> 2;                 // 8-byte vectors per input row
        for (int t = tid; t < area / 4; t += tps) {
          const int r = t / vpr, c = t - r * vpr;
          *reinterpret_cast<uint2 *>(bufA + r * ip + 4 * c) = __ldg(reinterpret_cast<const uint2 *>(in + tu * area) + t);
        }
      }
    }
    __syncthreads();
    if (active) fast_pass(bufA, bufB, sM1, f1, tid, tps);
    __syncthreads();
    if (active) fast_pass(bufB, bufA, sM2, f2, tid, tps);
    __syncthreads();
    if (active) {
      const int vpr = ow >> 2;
      for (int t = tid; t < area / 4; t += tps) {
        const int r = t / vpr, c = t - r * vpr;
        reinterpret_cast<uint2 *>(out + tu * area)[t] = *reinterpret_cast<const uint2 *>(bufA + r * op + 4 * c);
      }
    }
  }
}

int launch_transform(bool inverse, const TrPlan &p, const int16_t *in, int16_t *out, size_t n_tus,
                     cudaStream_t st)
{
  {
    const int rc = launch_transform_imma(inverse, p, in, out, n_tus, st);
    if (rc >= 0) return rc;
  }
  const int area = p.w * p.h;
  const int cl1 = inverse ? p.cutoff1 : p.n1, cl2 = inverse ? p.cutoff2 : p.n2;     // contraction lengths
  if (p.n2 && !(p.n1 & 3) && !(p.n2 & 3) && !(cl1 & 3) && !(cl2 & 3) && area >= 64 &&
      !(reinterpret_cast<uintptr_t>(in) & 15) && !(reinterpret_cast<uintptr_t>(out) & 15)) {
    // forward: output cutoff, full contraction; inverse: full output, contraction stops at the cutoff
    const FastPass f1{ p.n1, cl1, inverse ? p.n1 : p.cutoff1, p.line1, p.reduced1, p.shift1, inverse ? 1 : 0, 1 };
    const FastPass f2{ p.n2, cl2, inverse ? p.n2 : p.cutoff2, p.line2, p.reduced2, p.shift2, inverse ? 1 : 0, inverse ? 0 : 1 };
    const int items = (p.n1 / 4) * (p.line1 / 4);     // 4 x 4 output tiles per TU and pass (same for both passes)
    int slots = 256 / items;
    if (slots < 1) slots = 1;
    const int mx = p.w > p.h ? p.w : p.h, bsz = (mx + 4) * mx;
    while ((size_t)slots * 2 * bsz * sizeof(int16_t) > 40 * 1024) slots >>= 1;
    const size_t smem = ((p.n1 * (p.n1 + 4) + p.n2 * (p.n2 + 4) + 15) & ~15) + (size_t)slots * 2 * bsz * sizeof(int16_t);
    size_t grid = (n_tus + slots - 1) / slots;
    const size_t cap = (size_t)sm_count() * 8;
    if (grid > cap) grid = cap;
    if (inverse) transform_fast_kernel<true><<<(unsigned)grid, 256, smem, st>>>(in, out, n_tus, p, slots, f1, f2);
    else transform_fast_kernel<false><<<(unsigned)grid, 256, smem, st>>>(in, out, n_tus, p, slots, f1, f2);
    UVGB_LAUNCH_CHECK();
    return 0;
  }
  int slots = 256 / (area < 16 ? 16 : area);
  if (slots < 1) slots = 1;
  if (slots > 16) slots = 16;
  const size_t smem = ((p.n1 * p.n1 + p.n2 * p.n2 + 15) & ~15) + (size_t)slots * 2 * area * sizeof(int16_t);
  size_t grid = (n_tus + slots - 1) / slots;
  const size_t cap = (size_t)sm_count() * 8;
  if (grid > cap) grid = cap;
  if (inverse) transform_kernel<true><<<(unsigned)grid, 256, smem, st>>>(in, out, n_tus, p, slots);
  else transform_kernel<false><<<(unsigned)grid, 256, smem, st>>>(in, out, n_tus, p, slots);
  UVGB_LAUNCH_CHECK();
  return 0;
}

}  // namespace uvgb

using namespace uvgb;

extern "C" int uvg_b200_transform_dev(int inverse, int bitdepth, int w, int h, int type_hor,
                                      int type_ver, int skip_w, int skip_h, const int16_t *in,
                                      int16_t *out, size_t n_tus, void *stream)
{
  if (ensure_init()) return 1;
  if (n_tus == 0) return 0;
  if (bitdepth < 8 || bitdepth > 16) { set_error("bitdepth %d out of range", bitdepth); return 1; }
  if (skip_w < 0 || skip_w >= w + (w == 1) || skip_h < 0 || skip_h >= h + (h == 1)) { set_error("bad skip"); return 1; }
  TrPlan p;
  if (inverse ? make_inv_plan(p, bitdepth, w, h, type_hor, type_ver, skip_w, skip_h)
              : make_fwd_plan(p, bitdepth, w, h, type_hor, type_ver, skip_w, skip_h)) return 1;
  return launch_transform(inverse != 0, p, in, out, n_tus, static_cast<cudaStream_t>(stream));
}

extern "C" void uvg_b200_get_tr_type(int w, int h, int color, int cu_type, int tr_idx,
                                     int lfnst_idx, int cr_lfnst_idx, int isp_mode, int mts_type,
                                     int *hor, int *ver)
{
  int sw, sh;
  derive_tr(w, h, color, cu_type, tr_idx, lfnst_idx, cr_lfnst_idx, isp_mode, mts_type, hor, ver, &sw, &sh);
}
