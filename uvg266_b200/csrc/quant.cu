// quant.cu - scalar quantisation / dequantisation and the fused
// residual -> transform -> quant -> dequant -> inverse transform -> reconstruction kernel
// for sm_100a. Restates quant-generic.c: uvg_quant_generic :51-234 (incl. sign-bit hiding),
// uvg_dequant_generic :618-669 (flat lists), uvg_quantize_residual_generic :460-612,
// coeff_abs_sum_generic :671, fast_coeff_cost_generic :688.
#include "transform.cuh"
#include "quant.cuh"
#include <mutex>
#include <string.h>
#include <stdlib.h>

namespace uvgb {

int launch_transform(bool inverse, const TrPlan &p, const int16_t *in, int16_t *out, size_t n_tus,
                     cudaStream_t st);

// -------------------------------------------------------------------------------------
// scan tables (uvg_get_scan_order_table(SCAN_GROUP_4X4, SCAN_DIAG, ...), tables.c:2646):
// coefficient groups (uvg_g_log2_sbb_size, tables.c:13-24) in up-right diagonal order,
// diagonal order inside each group.
// -------------------------------------------------------------------------------------
static uint16_t h_scan[63 * 63];
static int h_scan_off[6][6];
static uint16_t *d_scan = nullptr;
static std::once_flag scan_once;
static cudaError_t scan_err = cudaSuccess;

static void cg_log2(int l2w, int l2h, int *cw, int *ch)
{
  if (l2w >= 2 && l2h >= 2) { *cw = 2; *ch = 2; return; }
  if (l2w == 0) { *cw = 0; *ch = l2h < 4 ? l2h : 4; return; }
  if (l2h == 0) { *ch = 0; *cw = l2w < 4 ? l2w : 4; return; }
  if (l2w == 1) { *cw = 1; *ch = (l2h <= 2) ? 1 : 3; return; }
  *ch = 1; *cw = (l2w <= 2) ? 1 : 3;
}

static int diag_fill(int w, int h, int *xs, int *ys)
{
  int n = 0;
  for (int s = 0; s < w + h - 1; ++s)
    for (int x = 0; x < w; ++x) {
      const int y = s - x;
      if (y >= 0 && y < h) { xs[n] = x; ys[n] = y; ++n; }
    }
  return n;
}

static void build_scan()
{
  int off = 0;
  static int gx[1024], gy[1024], ix[1024], iy[1024];
  for (int l2w = 0; l2w <= 5; ++l2w)
    for (int l2h = 0; l2h <= 5; ++l2h) {
      h_scan_off[l2w][l2h] = off;
      int cw, ch;
      cg_log2(l2w, l2h, &cw, &ch);
      const int w = 1 << l2w, gw = 1 << cw, gh = 1 << ch;
      const int ng = diag_fill(w >> cw, (1 << l2h) >> ch, gx, gy);
      const int ni = diag_fill(gw, gh, ix, iy);
      for (int g = 0; g < ng; ++g)
        for (int i = 0; i < ni; ++i)
          h_scan[off++] = (uint16_t)((gy[g] * gh + iy[i]) * w + gx[g] * gw + ix[i]);
    }
  scan_err = cudaMalloc(&d_scan, sizeof(h_scan));
  if (scan_err == cudaSuccess) scan_err = cudaMemcpy(d_scan, h_scan, sizeof(h_scan), cudaMemcpyHostToDevice);
}

int init_quant_tables()
{
  std::call_once(scan_once, build_scan);
  if (scan_err != cudaSuccess) { set_error("scan tables: %s", cudaGetErrorString(scan_err)); return 1; }
  return 0;
}

// -------------------------------------------------------------------------------------
// LFNST (transform.c:878-1200): the trained kernels are the encoder's own tables, handed in at run time
// -------------------------------------------------------------------------------------
static int8_t *d_lfnst8 = nullptr, *d_lfnst4 = nullptr;     // [4][2][16][48], [4][2][16][16]
static std::mutex lfnst_mu;

// uvg_lfnst_lut (lfnst_tables.h:51-54) as intervals of the extended mode index 0..94
static int lfnst_set(int m) { return m < 2 ? 0 : m <= 12 ? 1 : m <= 23 ? 2 : m <= 44 ? 3 : m <= 55 ? 2 : 1; }

// uvg_wide_angle_correction, intra.c:637-658
static int wide_angle_mode(int mode, int l2w, int l2h, int account_for_dc_planar)
{
  static const int shift[6] = { 0, 6, 10, 12, 14, 15 };
  if (l2w != l2h && mode > 1 && mode <= 66) {
    const int d = l2w > l2h ? l2w - l2h : l2h - l2w;
    if (l2w > l2h && mode < 2 + shift[d]) return mode + 65;
    if (l2h > l2w && mode > 66 - shift[d]) return mode - 65 - (account_for_dc_planar ? 2 : 0);
  }
  return mode;
}

// Position of LFNST scan entry j in a block of width w: up-right diagonal order inside a 4x4 group and over the
// 2x2 groups of the top-left 8x8 (uvg_coef_top_left_diag_scan_8x8 / the first group of uvg_g_sig_last_scan).
__device__ __forceinline__ int lfnst_scan_pos(int j, int w, int whge3)
{
  // x / y of the 16 diagonal positions, 2 bits each: dx = {0,0,1,0,1,2,0,1,2,3,1,2,3,2,3,3},
  // dy = {0,1,0,2,1,0,3,2,1,0,3,2,1,3,2,3}
  const unsigned DX = 0xfb9e4910u, DY = 0xedb1b184u;
  const int g = whge3 ? j >> 4 : 0, i = j & 15;
  const int gx = g >> 1, gy = g & 1;              // 4x4 groups of the 8x8 corner in the order (0,0), (0,1), (1,0), (1,1) as (x, y)
  return (gy * 4 + (int)((DY >> (2 * i)) & 3)) * w + gx * 4 + (int)((DX >> (2 * i)) & 3);
}

// index into the gathered input vector `in` (48 or 16 entries) of region coefficient (x, y) - the order in which
// uvg_fwd_lfnst reads the top-left region (row-major, or transposed; 8x8 regions are the L-shaped 48), :1030-1066
__device__ __forceinline__ int lfnst_in_index(int x, int y, int sb, int transpose)
{
  if (!transpose) return sb == 4 ? 4 * y + x : (y < 4 ? 8 * y + x : 32 + 4 * (y - 4) + x);
  if (sb == 4) return 4 * x + y;
  return x < 4 ? 8 * x + y : 32 + 4 * (x - 4) + y;
}

// -------------------------------------------------------------------------------------
// quantiser parameters (host)
// -------------------------------------------------------------------------------------
static const int16_t k_quant_scales[2][6] = { { 26214, 23302, 20560, 18396, 16384, 14564 },
                                              { 18396, 16384, 14564, 13107, 11651, 10280 } };  // scalinglist.c:91
static const int16_t k_inv_quant_scales[2][6] = { { 40, 45, 51, 57, 64, 72 },
                                                  { 57, 64, 72, 80, 90, 102 } };               // scalinglist.c:95

// uvg_get_scaled_qp, transform.c:150-165
static int scaled_qp(const uvg_b200_quant_cfg *c, int color)
{
  const int off = (c->bitdepth - 8) * 6;
  if (color == 0) return c->qp + off;
  int q = c->qp < -off ? -off : (c->qp > 57 ? 57 : c->qp);
  if (c->has_chroma_qp_map) return c->chroma_qp_map[c->qp & 127] + off;
  return q + off;
}

static int make_qparams(QParams &q, const uvg_b200_quant_cfg *c, int w, int h, int color,
                        int transform_skip, int lfnst_idx)
{
  if (!c) { set_error("quant cfg is NULL"); return 1; }
  if (c->bitdepth < 8 || c->bitdepth > 12) { set_error("quant: bitdepth %d unsupported", c->bitdepth); return 1; }
  auto pow2 = [](int v) { return v >= 1 && v <= 32 && (v & (v - 1)) == 0; };
  if (!pow2(w) || !pow2(h)) { set_error("quant: block %dx%d unsupported", w, h); return 1; }
  const int l2w = ilog2_host(w), l2h = ilog2_host(h);
  int qp_scaled = scaled_qp(c, color);
  if (transform_skip && qp_scaled < 4 + 6 * 2) qp_scaled = 4 + 6 * 2;   // MIN_QP_PRIME_TS
  if (qp_scaled < 0) { set_error("negative scaled qp"); return 1; }
  const int needs_scale = !transform_skip && ((l2w + l2h) % 2 == 1);
  const int transform_shift = 15 - c->bitdepth - ((l2w + l2h) >> 1) - needs_scale;       // :76
  q.w = w; q.h = h; q.area = w * h;
  q.lfnst_apply = 0; q.lfnst_transpose = 0; q.lfnst_sb = 0; q.lfnst_zero_out = 0; q.lfnst_whge3 = 0; q.lfnst_mat = nullptr;
  q.q_bits = 14 + qp_scaled / 6 + (transform_skip ? 0 : transform_shift);                // :77
  if (q.q_bits < 9 || q.q_bits > 30) { set_error("q_bits %d out of range", q.q_bits); return 1; }
  q.add = (c->slice_intra ? 171 : 85) << (q.q_bits - 9);                                 // :78
  q.q_bits8 = q.q_bits - 8;
  q.coeff_default = k_quant_scales[needs_scale][qp_scaled % 6];                          // :81
  q.coeff_list = k_quant_scales[0][qp_scaled % 6];   // flat quant_coeff[] (scalinglist.c:416)
  q.signhide = c->signhide;
  q.lfnst_idx = lfnst_idx;
  q.max_lfnst = ((w == 4 && h == 4) || (w == 8 && h == 8)) ? 8 : 16;                    // :103
  if (lfnst_idx && (w < 4 || h < 4)) { set_error("lfnst needs a >=4x4 block"); return 1; }
  if (c->signhide && q.area < 16) { set_error("sign hiding needs >= 16 coefficients"); return 1; }
  q.scan = d_scan + h_scan_off[l2w][l2h];
  // dequant :629-637, :660-661
  const int ts_deq = 15 - c->bitdepth - ((l2w + l2h) >> 1);
  q.dq_shift = 20 - 14 - (transform_skip ? 0 : ts_deq - needs_scale);
  if (q.dq_shift < 1) { set_error("dequant shift %d unsupported for %dx%d", q.dq_shift, w, h); return 1; }
  q.dq_scale = k_inv_quant_scales[needs_scale][qp_scaled % 6] << (qp_scaled / 6);
  q.dq_add = 1 << (q.dq_shift - 1);
  q.pixel_max = (1 << c->bitdepth) - 1;
  q.ssd_shift = 2 * (c->bitdepth - 8);
  return 0;
}

// -------------------------------------------------------------------------------------
// device building blocks (operate on one TU held in shared memory; `tid`/`nt` index the
// threads cooperating on that TU)
// -------------------------------------------------------------------------------------
__device__ __forceinline__ int quant_level(int c, int coeff, const QParams &q)
{
  const unsigned a = (unsigned)abs(c);
  // |c| <= 32768 and coeff <= 26214: the product + add stays below 2^31
  int level = (int)((a * (unsigned)coeff + (unsigned)q.add) >> q.q_bits);
  return level;
}

// level loop of uvg_quant_generic (:86-121); returns this thread's partial ac_sum
__device__ __forceinline__ unsigned quant_elems(const int16_t *coef, int16_t *qc, const QParams &q,
                                                int tid, int nt)
{
  unsigned ac = 0;
  if (q.lfnst_idx == 0) {
    for (int n = tid; n < q.area; n += nt) {
      const int c = coef[n];
      int level = quant_level(c, q.coeff_default, q);
      ac += (unsigned)level;
      level = c < 0 ? -level : level;
      qc[n] = (int16_t)clip3(-32768, 32767, level);
    }
  } else {
    // memset(q_coef, 0) :104; the kept coefficients are written by quant_lfnst_elems after
    // a barrier
    for (int n = tid; n < q.area; n += nt) qc[n] = 0;
  }
  return ac;
}

__device__ __forceinline__ unsigned quant_lfnst_elems(const int16_t *coef, int16_t *qc,
                                                      const QParams &q, int tid, int nt)
{
  unsigned ac = 0;
  for (int n = tid; n < q.max_lfnst; n += nt) {
    const int idx = q.scan[n];
    const int c = coef[idx];
    int level = quant_level(c, q.coeff_list, q);   // always the list table, :114
    ac += (unsigned)level;
    level = c < 0 ? -level : level;
    qc[idx] = (int16_t)clip3(-32768, 32767, level);
  }
  return ac;
}

__device__ __forceinline__ int delta_u_of(int c, const QParams &q)
{
  const unsigned a = (unsigned)abs(c);
  const unsigned prod = a * (unsigned)q.coeff_list;        // list table even when lists are off (:132,:144)
  const int level = (int)((prod + (unsigned)q.add) >> q.q_bits);
  return ((int)prod - (level << q.q_bits)) >> q.q_bits8;
}

// Sign-bit hiding for coefficient group `cg` (:151-233). `is_last_cg`: this is the first
// group, scanning from the end, that holds a non-zero level.
__device__ void signhide_cg(const int16_t *coef, int16_t *qc, const QParams &q, int cg, bool is_last_cg)
{
  const uint16_t *scan = q.scan + (cg << 4);
  int first_nz = 16, last_nz = -1;
  for (int n = 15; n >= 0; --n) if (qc[scan[n]]) { last_nz = n; break; }
  if (last_nz < 0) return;
  for (int n = 0; n < 16; ++n) if (qc[scan[n]]) { first_nz = n; break; }
  if (last_nz - first_nz < 4) return;
  int abssum = 0;
  for (int n = first_nz; n <= last_nz; ++n) abssum += qc[scan[n]];
  const int signbit = qc[scan[first_nz]] > 0 ? 0 : 1;
  if (signbit == (abssum & 1)) return;
  int min_cost = 0x7fffffff, cur_cost = 0x7fffffff, min_pos = -1;
  int final_change = 0, cur_change = 0;
  for (int n = is_last_cg ? last_nz : 15; n >= 0; --n) {
    const int pos = scan[n];
    const int qv = qc[pos];
    if (qv != 0) {
      const int du = delta_u_of(coef[pos], q);
      if (du > 0) { cur_cost = -du; cur_change = 1; }
      else if (n == first_nz && abs(qv) == 1) { cur_cost = 0x7fffffff; }
      else { cur_cost = du; cur_change = -1; }
    } else if (n < first_nz && ((coef[pos] >= 0) ? 0 : 1) != signbit) {
      cur_cost = 0x7fffffff;
    } else {
      cur_cost = -delta_u_of(coef[pos], q); cur_change = 1;
    }
    if (cur_cost < min_cost) { min_cost = cur_cost; final_change = cur_change; min_pos = pos; }
  }
  if (qc[min_pos] == 32767 || qc[min_pos] == -32768) final_change = -1;
  if (coef[min_pos] >= 0) qc[min_pos] = (int16_t)(qc[min_pos] + final_change);
  else qc[min_pos] = (int16_t)(qc[min_pos] - final_change);
}

// dequant, flat branch (:660-668)
__device__ __forceinline__ void dequant_elems(const int16_t *qc, int16_t *coef, const QParams &q,
                                              int tid, int nt)
{
  for (int n = tid; n < q.area; n += nt) {
    const int v = ((int)qc[n] * q.dq_scale + q.dq_add) >> q.dq_shift;
    coef[n] = (int16_t)clip3(-32768, 32767, v);
  }
}

// Whole quant step for the TUs of a CTA, shared by the standalone and the fused kernel.
// s_coef / s_q: per-slot buffers; s_acc[slot] / s_cgnz[slot*64..]: per-slot scratch.
// All threads of the CTA must call it (contains __syncthreads).
__device__ __forceinline__ void quant_step(const int16_t *s_coef, int16_t *s_q, const QParams &q,
                                           unsigned *s_acc, unsigned char *s_cgnz, int slot, int tid,
                                           int nt, bool active)
{
  if (tid == 0) s_acc[slot] = 0;
  __syncthreads();
  if (active) {
    unsigned ac = quant_elems(s_coef, s_q, q, tid, nt);
    if (q.lfnst_idx == 0 && ac) atomicAdd(&s_acc[slot], ac);
  }
  if (q.lfnst_idx) {
    __syncthreads();
    if (active) {
      unsigned ac = quant_lfnst_elems(s_coef, s_q, q, tid, nt);
      if (ac) atomicAdd(&s_acc[slot], ac);
    }
  }
  __syncthreads();
  if (q.signhide) {
    const int ncg = q.area >> 4;
    unsigned char *nz = s_cgnz + slot * 64;
    if (active && s_acc[slot] >= 2) {
      for (int cg = tid; cg < ncg; cg += nt) {
        bool any = false;
        for (int n = 0; n < 16; ++n) any |= s_q[q.scan[(cg << 4) + n]] != 0;
        nz[cg] = any;
      }
    }
    __syncthreads();
    if (active && s_acc[slot] >= 2) {
      for (int cg = tid; cg < ncg; cg += nt) {
        if (!nz[cg]) continue;
        bool higher = false;
        for (int c2 = cg + 1; c2 < ncg; ++c2) higher |= nz[c2] != 0;
        signhide_cg(s_coef, s_q, q, cg, !higher);
      }
    }
    __syncthreads();
  }
}

// -------------------------------------------------------------------------------------
// standalone quant / dequant kernels
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
quant_kernel(const int16_t *__restrict__ coef, int16_t *__restrict__ qcoef, size_t n_tus, QParams q,
             int slots)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int16_t *sbuf = reinterpret_cast<int16_t *>(smem_raw);
  unsigned *s_acc = reinterpret_cast<unsigned *>(sbuf + (size_t)slots * 2 * q.area);
  unsigned char *s_cgnz = reinterpret_cast<unsigned char *>(s_acc + slots);
  const int tps = blockDim.x / slots;
  const int slot = threadIdx.x / tps, tid = threadIdx.x - slot * tps;
  int16_t *s_coef = sbuf + (size_t)slot * 2 * q.area, *s_q = s_coef + q.area;
  for (size_t base = (size_t)blockIdx.x * slots; base < n_tus; base += (size_t)gridDim.x * slots) {
    const size_t tu = base + slot;
    const bool active = tu < n_tus;
    __syncthreads();
    if (active) for (int t = tid; t < q.area; t += tps) s_coef[t] = coef[tu * q.area + t];
    __syncthreads();
    quant_step(s_coef, s_q, q, s_acc, s_cgnz, slot, tid, tps, active);
    if (active) for (int t = tid; t < q.area; t += tps) qcoef[tu * q.area + t] = s_q[t];
  }
}

// Without sign hiding and LFNST, quant (:86-121) and dequant (:660-668) are element-wise: 8 coefficients
// (one 16-byte vector) per thread and step, no shared memory. `total` is a multiple of 8 (area >= 16 / 8).
template <bool DEQUANT>
__global__ void __launch_bounds__(256)
quant_flat_kernel(const int16_t *__restrict__ in, int16_t *__restrict__ out, size_t total8, QParams q)
{
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total8; t += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in) + t);
    const uint32_t w[4] = { v.x, v.y, v.z, v.w };
    uint32_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int o[2];
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        const int c = hlf ? (int)w[i] >> 16 : (int)(short)(w[i] & 0xffff);
        if (DEQUANT) {
          o[hlf] = clip3(-32768, 32767, (c * q.dq_scale + q.dq_add) >> q.dq_shift);
        } else {
          const int level = quant_level(c, q.coeff_default, q);
          o[hlf] = clip3(-32768, 32767, c < 0 ? -level : level);
        }
      }
      r[i] = ((uint32_t)o[0] & 0xffff) | ((uint32_t)o[1] << 16);
    }
    reinterpret_cast<uint4 *>(out)[t] = make_uint4(r[0], r[1], r[2], r[3]);
  }
}

__global__ void __launch_bounds__(256)
dequant_kernel(const int16_t *__restrict__ qcoef, int16_t *__restrict__ coef, size_t total, QParams q)
{
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (size_t)gridDim.x * blockDim.x) {
    const int v = ((int)qcoef[t] * q.dq_scale + q.dq_add) >> q.dq_shift;
    coef[t] = (int16_t)clip3(-32768, 32767, v);
  }
}

// -------------------------------------------------------------------------------------
// fused quantize_residual
// -------------------------------------------------------------------------------------
template <typename Pix>
__global__ void __launch_bounds__(256)
quantize_residual_kernel(FusedArgs a, size_t n_tus, QParams q, TrPlan fp, TrPlan ip, int slots)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // matrices: forward pair transposed, inverse pair row-major
  int8_t *sF1 = reinterpret_cast<int8_t *>(smem_raw);
  int8_t *sF2 = sF1 + fp.n1 * fp.n1;
  int8_t *sI1 = sF2 + fp.n2 * fp.n2;
  int8_t *sI2 = sI1 + ip.n1 * ip.n1;
  const int mat_bytes = (fp.n1 * fp.n1 + fp.n2 * fp.n2 + ip.n1 * ip.n1 + ip.n2 * ip.n2 + 15) & ~15;
  int16_t *sbuf = reinterpret_cast<int16_t *>(smem_raw + mat_bytes);
  unsigned *s_acc = reinterpret_cast<unsigned *>(sbuf + (size_t)slots * 3 * q.area);
  unsigned *s_red = s_acc + slots;                       // [slots][3]: nonzero, ssd, cost
  unsigned char *s_cgnz = reinterpret_cast<unsigned char *>(s_red + slots * 3);
  int16_t *s_lfnst = reinterpret_cast<int16_t *>(s_cgnz + slots * 64);        // [slots][64]: LFNST input / output vectors
  if (!a.use_trskip) {
    for (int t = threadIdx.x; t < fp.n1 * fp.n1; t += blockDim.x) { const int k = t / fp.n1, i = t - k * fp.n1; sF1[i * fp.n1 + k] = fp.t1[t]; }
    for (int t = threadIdx.x; t < fp.n2 * fp.n2; t += blockDim.x) { const int k = t / fp.n2, i = t - k * fp.n2; sF2[i * fp.n2 + k] = fp.t2[t]; }
    for (int t = threadIdx.x; t < ip.n1 * ip.n1; t += blockDim.x) sI1[t] = ip.t1[t];
    for (int t = threadIdx.x; t < ip.n2 * ip.n2; t += blockDim.x) sI2[t] = ip.t2[t];
  }
  const int tps = blockDim.x / slots;
  const int slot = threadIdx.x / tps, tid = threadIdx.x - slot * tps;
  int16_t *bufA = sbuf + (size_t)slot * 3 * q.area, *bufB = bufA + q.area, *bufQ = bufB + q.area;
  const int w = q.w, area = q.area;
  for (size_t base = (size_t)blockIdx.x * slots; base < n_tus; base += (size_t)gridDim.x * slots) {
    const size_t tu = base + slot;
    const bool active = tu < n_tus;
    const Pix *pref = nullptr, *ppred = nullptr; Pix *prec = nullptr;
    if (active) {
      pref = static_cast<const Pix *>(a.ref.base) + a.ref.offset(tu);
      ppred = static_cast<const Pix *>(a.pred.base) + a.pred.offset(tu);
      prec = const_cast<Pix *>(static_cast<const Pix *>(a.rec.base)) + a.rec.offset(tu);
    }
    __syncthreads();
    if (tid < 3) s_red[slot * 3 + tid] = 0;
    // 1. residual (generate_residual :480)
    const Pix *pref2 = nullptr, *ppred2 = nullptr; Pix *prec2 = nullptr;
    if (active && a.jccr) {
      pref2 = static_cast<const Pix *>(a.ref2.base) + a.ref2.offset(tu);
      ppred2 = static_cast<const Pix *>(a.pred2.base) + a.pred2.offset(tu);
      prec2 = const_cast<Pix *>(static_cast<const Pix *>(a.rec2.base)) + a.rec2.offset(tu);
    }
    if (active)
      for (int t = tid; t < area; t += tps) {
        const int y = t / w, x = t - y * w;
        const int cb = (int)pref[(size_t)y * a.ref.stride + x] - (int)ppred[(size_t)y * a.pred.stride + x];
        if (!a.jccr) { bufA[t] = (int16_t)cb; continue; }
        // joint residual, quant-generic.c:275-309 (C division: truncation toward zero)
        const int cr = (int)pref2[(size_t)y * a.ref2.stride + x] - (int)ppred2[(size_t)y * a.pred2.stride + x];
        int v;
        switch (a.jccr) {
          case 2: v = (4 * cb + 2 * cr) / 5; break;
          case -2: v = (4 * cb - 2 * cr) / 5; break;
          case 3: v = (cb + cr) / 2; break;
          case -3: v = (cb - cr) / 2; break;
          case 1: v = (4 * cr + 2 * cb) / 5; break;
          default: v = (4 * cr - 2 * cb) / 5; break;      // -1
        }
        bufA[t] = (int16_t)v;
      }
    __syncthreads();
    // 2. forward transform (:495-501) -> coefficients in bufA
    if (!a.use_trskip) {
      if (active) fwd_pass(bufA, bufB, sF1, fp.n1, fp.line1, fp.reduced1, fp.cutoff1, fp.shift1, tid, tps);
      __syncthreads();
      if (fp.n2) {
        if (active) fwd_pass(bufB, bufA, sF2, fp.n2, fp.line2, fp.reduced2, fp.cutoff2, fp.shift2, tid, tps);
      } else {
        if (active) for (int t = tid; t < area; t += tps) bufA[t] = bufB[t];
      }
      __syncthreads();
    }
    // 2b. forward LFNST on the low-frequency corner (uvg_fwd_lfnst, transform.c:965-1083): gather the 4x4 / L-shaped
    // 8x8 region (transposed for modes > 34), 16 x 16|48 matrix-vector product, scatter in scan order
    if (q.lfnst_apply) {
      int16_t *lf = s_lfnst + slot * 64;
      const int sb = q.lfnst_sb, tr = sb > 4 ? 48 : 16;
      if (active)
        for (int e = tid; e < sb * sb; e += tps) {
          const int y = e / sb, x = e - y * sb;
          if (sb == 8 && y >= 4 && x >= 4) continue;
          lf[lfnst_in_index(x, y, sb, q.lfnst_transpose)] = bufA[y * w + x];
        }
      __syncthreads();
      int16_t o[3] = { 0, 0, 0 };
      if (active)
        for (int j = tid, r = 0; j < tr; j += tps, ++r) {
          if (j < q.lfnst_zero_out) {
            int acc = 0;
            for (int i = 0; i < tr; ++i) acc += (int)lf[i] * (int)q.lfnst_mat[j * tr + i];
            o[r] = (int16_t)((acc + 64) >> 7);
          }
        }
      __syncthreads();
      if (active)
        for (int j = tid, r = 0; j < tr; j += tps, ++r) bufA[lfnst_scan_pos(j, w, q.lfnst_whge3)] = o[r];
      __syncthreads();
    }
    // 3. quant (+ sign hiding) :536-539
    quant_step(bufA, bufQ, q, s_acc, s_cgnz, slot, tid, tps, active);
    // 4. coefficients out, has_coeffs (:542-550), fast coefficient cost
    if (active) {
      unsigned nzf = 0, cst = 0;
      for (int t = tid; t < area; t += tps) {
        const int16_t v = bufQ[t];
        a.coeff_out[tu * area + t] = v;
        nzf |= (v != 0);
        const unsigned av = min((unsigned)abs((int)v), 3u);
        cst += av == 0 ? a.w0 : av == 1 ? a.w1 : av == 2 ? a.w2 : a.w3;
      }
      if (nzf) atomicOr(&s_red[slot * 3 + 0], 1u);
      if (a.cost && cst) atomicAdd(&s_red[slot * 3 + 2], cst);
    }
    __syncthreads();
    const bool has = active && s_red[slot * 3 + 0] != 0;
    const bool recon = has && !a.early_skip;
    // 5. dequant + inverse transform (:553-570) -> residual' in bufA
    if (recon) dequant_elems(bufQ, a.use_trskip ? bufA : bufB, q, tid, tps);
    __syncthreads();
    // 5b. inverse LFNST (uvg_inv_lfnst, transform.c:1104-1200) on the dequantised coefficients (bufB)
    if (q.lfnst_apply) {
      int16_t *lf = s_lfnst + slot * 64;
      const int sb = q.lfnst_sb, tr = sb > 4 ? 48 : 16;
      if (recon) for (int j = tid; j < 16; j += tps) lf[48 + j] = bufB[lfnst_scan_pos(j, w, q.lfnst_whge3)];
      __syncthreads();
      if (recon)
        for (int j = tid; j < tr; j += tps) {
          int acc = 0;
          for (int i = 0; i < q.lfnst_zero_out; ++i) acc += (int)lf[48 + i] * (int)q.lfnst_mat[i * tr + j];
          lf[j] = (int16_t)((acc + 64) >> 7);
        }
      __syncthreads();
      if (recon)
        for (int e = tid; e < sb * sb; e += tps) {
          const int y = e / sb, x = e - y * sb;
          if (sb == 8 && y >= 4 && x >= 4) continue;
          bufB[y * w + x] = lf[lfnst_in_index(x, y, sb, q.lfnst_transpose)];
        }
      __syncthreads();
    }
    if (!a.use_trskip) {
      // bufB holds dequantised coefficients
      if (recon) inv_pass(bufB, ip.n2 ? bufQ : bufA, sI1, ip.n1, ip.line1, ip.reduced1, ip.cutoff1, ip.shift1, tid, tps);
      __syncthreads();
      if (ip.n2) {
        if (recon) inv_pass(bufQ, bufA, sI2, ip.n2, ip.line2, ip.reduced2, ip.cutoff2, ip.shift2, tid, tps);
        __syncthreads();
      }
    }
    // 6. reconstruction (:591-611) + SSD against the source
    if (active && a.jccr) {
      // split the joint residual back (quant-generic.c:388-429), both planes; without coefficients both
      // reconstructions are copies of the predictions (:431-441)
      for (int t = tid; t < area; t += tps) {
        const int y = t / w, x = t - y * w;
        const int pu = (int)ppred[(size_t)y * a.pred.stride + x], pv2 = (int)ppred2[(size_t)y * a.pred2.stride + x];
        int ru = pu, rv2 = pv2;
        if (recon) {
          const int c = bufA[t];
          int16_t ur, vr;
          switch (a.jccr) {
            case 2: ur = (int16_t)c; vr = (int16_t)(c >> 1); break;
            case -2: ur = (int16_t)c; vr = (int16_t)(-c >> 1); break;
            case 3: ur = (int16_t)c; vr = (int16_t)c; break;
            case -3: ur = (int16_t)c; vr = (int16_t)(-c); break;
            case 1: ur = (int16_t)(c >> 1); vr = (int16_t)c; break;
            default: ur = (int16_t)(-c >> 1); vr = (int16_t)c; break;   // -1
          }
          ru = clip3(0, q.pixel_max, (int)(int16_t)(ur + pu));
          rv2 = clip3(0, q.pixel_max, (int)(int16_t)(vr + pv2));
        }
        prec[(size_t)y * a.rec.stride + x] = (Pix)ru;
        prec2[(size_t)y * a.rec2.stride + x] = (Pix)rv2;
      }
    } else if (active) {
      unsigned ssd = 0;
      for (int t = tid; t < area; t += tps) {
        const int y = t / w, x = t - y * w;
        const int pv = (int)ppred[(size_t)y * a.pred.stride + x];
        int rv = pv;
        if (recon) rv = clip3(0, q.pixel_max, (int)(short)(bufA[t] + pv));
        if (recon || !a.rec_is_pred) prec[(size_t)y * a.rec.stride + x] = (Pix)rv;
        const int d = (int)pref[(size_t)y * a.ref.stride + x] - rv;
        ssd += (unsigned)(d * d);
      }
      if (a.ssd && ssd) atomicAdd(&s_red[slot * 3 + 1], ssd);
    }
    __syncthreads();
    if (active && tid == 0) {
      a.has_coeffs[tu] = has ? 1 : 0;
      if (a.ssd) a.ssd[tu] = (uint32_t)((int32_t)s_red[slot * 3 + 1] >> q.ssd_shift);
      if (a.cost) a.cost[tu] = (s_red[slot * 3 + 2] + 128u) >> 8;
    }
  }
}

// -------------------------------------------------------------------------------------
// coefficient sums
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
coeff_sum_kernel(const int16_t *__restrict__ coeffs, size_t len, uint32_t *__restrict__ out, size_t n_tus,
                 int mode, unsigned w0, unsigned w1, unsigned w2, unsigned w3)
{
  // one warp per block of `len` coefficients
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_tus) return;
  const int16_t *p = coeffs + warp * len;
  unsigned s = 0;
  for (size_t i = lane; i < len; i += 32) {
    const unsigned av = (unsigned)abs((int)p[i]);
    if (mode == 0) s += av;
    else { const unsigned c = min(av, 3u); s += c == 0 ? w0 : c == 1 ? w1 : c == 2 ? w2 : w3; }
  }
  s = (unsigned)warp_sum((int)s);
  if (lane == 0) out[warp] = mode == 0 ? s : (s + 128u) >> 8;
}

static int slots_for(int area)
{
  int slots = 256 / (area < 16 ? 16 : area);
  if (slots < 1) slots = 1;
  if (slots > 16) slots = 16;
  return slots;
}

}  // namespace uvgb

using namespace uvgb;

extern "C" int uvg_b200_quant_dev(const uvg_b200_quant_cfg *cfg, const int16_t *coef, int16_t *q_coef,
                                  int w, int h, int color, int transform_skip, int lfnst_idx,
                                  size_t n_tus, void *stream)
{
  if (ensure_init()) return 1;
  if (n_tus == 0) return 0;
  QParams q;
  if (make_qparams(q, cfg, w, h, color, transform_skip, lfnst_idx)) return 1;
  {
    const size_t total = n_tus * (size_t)q.area;
    if (!q.signhide && !q.lfnst_idx && total % 8 == 0 && ((reinterpret_cast<uintptr_t>(coef) | reinterpret_cast<uintptr_t>(q_coef)) & 15) == 0) {
      size_t grid = (total / 8 + 255) / 256;
      const size_t cap = (size_t)sm_count() * 16;
      if (grid > cap) grid = cap;
      quant_flat_kernel<false><<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(coef, q_coef, total / 8, q);
      UVGB_LAUNCH_CHECK();
      return 0;
    }
  }
  const int slots = slots_for(q.area);
  const size_t smem = (size_t)slots * 2 * q.area * sizeof(int16_t) + slots * sizeof(unsigned) + slots * 64;
  size_t grid = (n_tus + slots - 1) / slots;
  const size_t cap = (size_t)sm_count() * 8;
  if (grid > cap) grid = cap;
  quant_kernel<<<(unsigned)grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(coef, q_coef, n_tus, q, slots);
  UVGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int uvg_b200_dequant_dev(const uvg_b200_quant_cfg *cfg, const int16_t *q_coef, int16_t *coef,
                                    int w, int h, int color, int transform_skip, size_t n_tus,
                                    void *stream)
{
  if (ensure_init()) return 1;
  if (n_tus == 0) return 0;
  QParams q;
  uvg_b200_quant_cfg c = *cfg; c.signhide = 0;
  if (make_qparams(q, &c, w, h, color, transform_skip, 0)) return 1;
  const size_t total = n_tus * (size_t)q.area;
  if (total % 8 == 0 && ((reinterpret_cast<uintptr_t>(coef) | reinterpret_cast<uintptr_t>(q_coef)) & 15) == 0) {
    size_t grid8 = (total / 8 + 255) / 256;
    const size_t cap8 = (size_t)sm_count() * 16;
    if (grid8 > cap8) grid8 = cap8;
    quant_flat_kernel<true><<<(unsigned)grid8, 256, 0, static_cast<cudaStream_t>(stream)>>>(q_coef, coef, total / 8, q);
    UVGB_LAUNCH_CHECK();
    return 0;
  }
  size_t grid = (total + 255) / 256;
  const size_t cap = (size_t)sm_count() * 16;
  if (grid > cap) grid = cap;
  dequant_kernel<<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(q_coef, coef, total, q);
  UVGB_LAUNCH_CHECK();
  return 0;
}

// Shared by quantize_residual and quant_cbcr_residual: plans, parameters, kernel choice.
// cr_* == nullptr: single plane. Otherwise the joint Cb-Cr mode of the general kernel (jccr != 0).
static int run_quantize_residual(const uvg_b200_quant_cfg *cfg, int cu_type, int tr_idx, int w, int h, int color,
                                 int use_trskip, int early_skip, const uvg_b200_blocks *ref,
                                 const uvg_b200_blocks *pred, const uvg_b200_blocks *rec,
                                 const uvg_b200_blocks *cr_ref, const uvg_b200_blocks *cr_pred,
                                 const uvg_b200_blocks *cr_rec, int jccr, int16_t *coeff_out, uint8_t *has_coeffs,
                                 uint32_t *ssd, uint32_t *cost, uint64_t cost_weights, size_t n_tus, void *stream)
{
  QParams q;
  // quant is called with transform_skip = (tr_idx == MTS_SKIP && luma), quant-generic.c:538
  const int ts_flag = (tr_idx == 1 && color == 0) ? 1 : 0;
  // quant is told the LFNST index of this colour / tree even where uvg_fwd_lfnst's own gate keeps the transform
  // off (quant-generic.c:503, :538): only the first 8 / 16 scan positions survive
  if (make_qparams(q, cfg, w, h, color, ts_flag, jccr ? 0 : cfg->lfnst_index)) return 1;
  if (cfg->lfnst_apply) {
    if (use_trskip || jccr || cu_type != 1 || !cfg->lfnst_index || cfg->lfnst_index > 2) { set_error("quantize_residual: LFNST with transform skip / JCCR / inter CU is not covered"); return 1; }
    const int8_t *t8, *t4;
    { std::lock_guard<std::mutex> lk(lfnst_mu); t8 = d_lfnst8; t4 = d_lfnst4; }
    if (!t8 || !t4) { set_error("quantize_residual: LFNST kernels not set (uvg_b200_set_lfnst_tables)"); return 1; }
    int m = wide_angle_mode(cfg->lfnst_mode, cfg->lfnst_wa_log2_w, cfg->lfnst_wa_log2_h, 1);
    m = m < 0 ? m + 14 + 67 : m >= 67 ? m + 14 : m;                    // get_lfnst_intra_mode, transform.c:917-936
    if (m < 0 || m > 94) { set_error("quantize_residual: LFNST mode %d out of range", cfg->lfnst_mode); return 1; }
    q.lfnst_apply = 1;
    q.lfnst_transpose = (m >= 67 && m >= 81) || (m < 67 && m > 34);    // get_transpose_flag, :938-942
    q.lfnst_whge3 = w >= 8 && h >= 8;
    q.lfnst_sb = q.lfnst_whge3 ? 8 : 4;
    q.lfnst_zero_out = ((w == 4 && h == 4) || (w == 8 && h == 8)) ? 8 : 16;
    const size_t sel = (size_t)(lfnst_set(m) * 2 + cfg->lfnst_index - 1) * 16;
    q.lfnst_mat = q.lfnst_sb > 4 ? t8 + sel * 48 : t4 + sel * 16;
  }
  TrPlan fp{}, ip{};
  int th = UVG_B200_DCT2, tv = UVG_B200_DCT2, sw = 0, sh = 0;
  if (!use_trskip) {
    // uvg_transform2d (transform.c:256): mts_dct when cfg.mts or non-square, else dct_NxN -
    // both are this matrix product; with mts off the types are DCT2/DCT2.
    derive_tr(w, h, color, cu_type, tr_idx, jccr ? 0 : cfg->cu_lfnst_idx, jccr ? 0 : cfg->cu_cr_lfnst_idx, cfg->isp_mode, cfg->mts, &th, &tv, &sw, &sh);
    if (make_fwd_plan(fp, cfg->bitdepth, w, h, th, tv, sw, sh)) return 1;
    if (make_inv_plan(ip, cfg->bitdepth, w, h, th, tv, sw, sh)) return 1;
  }
  FusedArgs a;
  a.ref = to_blocks(ref); a.pred = to_blocks(pred); a.rec = to_blocks(rec);
  a.coeff_out = coeff_out; a.has_coeffs = has_coeffs; a.ssd = ssd; a.cost = cost;
  a.w0 = (unsigned)(cost_weights & 0xffff); a.w1 = (unsigned)((cost_weights >> 16) & 0xffff);
  a.w2 = (unsigned)((cost_weights >> 32) & 0xffff); a.w3 = (unsigned)((cost_weights >> 48) & 0xffff);
  a.use_trskip = use_trskip; a.early_skip = early_skip;
  a.rec_is_pred = (rec->base == pred->base && rec->stride == pred->stride &&
                   rec->col_pitch == pred->col_pitch && rec->row_pitch == pred->row_pitch &&
                   rec->blocks_per_row == pred->blocks_per_row);
  a.jccr = jccr;
  if (jccr) { a.ref2 = to_blocks(cr_ref); a.pred2 = to_blocks(cr_pred); a.rec2 = to_blocks(cr_rec); }
  else { a.ref2 = a.ref; a.pred2 = a.pred; a.rec2 = a.rec; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // shape-specialised kernels first: registers-only for 4x4/8x8, tensor cores for 16x16/32x32
  const bool lfnst_involved = !jccr && (cfg->lfnst_apply || cfg->lfnst_index || cfg->cu_lfnst_idx || cfg->cu_cr_lfnst_idx);
  if (!jccr && !lfnst_involved) {
    // UVG266_B200_FUSED = imma | small | hmma | generic restricts the choice (A/B measurements)
    static const char *force = getenv("UVG266_B200_FUSED");
    int rc = -1;
    if (!force || !strcmp(force, "imma")) rc = launch_fused_imma(a, n_tus, q, cfg->bitdepth, th, tv, sw, sh, st);
    if (rc >= 0) return rc;
    if (!force || !strcmp(force, "small") || (!strcmp(force, "imma") && q.w == 4)) rc = launch_fused_small(a, n_tus, q, cfg->bitdepth, th, tv, st);
    if (rc >= 0) return rc;
    if (!force || !strcmp(force, "hmma")) rc = launch_fused_mma(a, n_tus, q, cfg->bitdepth, th, tv, sw, sh, st);
    if (rc >= 0) return rc;
  }
  const int slots = slots_for(q.area);
  const int mat_bytes = (fp.n1 * fp.n1 + fp.n2 * fp.n2 + ip.n1 * ip.n1 + ip.n2 * ip.n2 + 15) & ~15;
  const size_t smem = mat_bytes + (size_t)slots * 3 * q.area * sizeof(int16_t) + slots * 4 * sizeof(unsigned) + slots * 64 +
                      (size_t)slots * 64 * sizeof(int16_t);
  size_t grid = (n_tus + slots - 1) / slots;
  const size_t cap = (size_t)sm_count() * 8;
  if (grid > cap) grid = cap;
  if (cfg->bitdepth == 8) quantize_residual_kernel<uint8_t><<<(unsigned)grid, 256, smem, st>>>(a, n_tus, q, fp, ip, slots);
  else quantize_residual_kernel<uint16_t><<<(unsigned)grid, 256, smem, st>>>(a, n_tus, q, fp, ip, slots);
  UVGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int uvg_b200_quantize_residual_dev(const uvg_b200_quant_cfg *cfg, int cu_type, int tr_idx,
                                              int w, int h, int color, int use_trskip, int early_skip,
                                              const uvg_b200_blocks *ref, const uvg_b200_blocks *pred,
                                              const uvg_b200_blocks *rec, int16_t *coeff_out,
                                              uint8_t *has_coeffs, uint32_t *ssd, uint32_t *cost,
                                              uint64_t cost_weights, size_t n_tus, void *stream)
{
  if (ensure_init()) return 1;
  if (n_tus == 0) return 0;
  if (!cfg || !ref || !pred || !rec || !coeff_out || !has_coeffs) { set_error("quantize_residual: NULL argument"); return 1; }
  return run_quantize_residual(cfg, cu_type, tr_idx, w, h, color, use_trskip, early_skip, ref, pred, rec, nullptr,
                               nullptr, nullptr, 0, coeff_out, has_coeffs, ssd, cost, cost_weights, n_tus, stream);
}

// uvg_quant_cbcr_residual_generic (quant-generic.c:241-442) for the plain chain (no LFNST / RDOQ / dep-quant):
// joint residual of the two chroma planes -> transform -> quant -> has_coeffs -> dequant -> inverse transform ->
// split -> both reconstructions. The coded plane is V when joint_cb_cr == 1, else U (:311, :337).
extern "C" int uvg_b200_quant_cbcr_residual_dev(const uvg_b200_quant_cfg *cfg, int cu_type, int tr_idx, int w, int h,
                                                int joint_cb_cr, int jccr_sign, int early_skip,
                                                const uvg_b200_blocks *u_ref, const uvg_b200_blocks *v_ref,
                                                const uvg_b200_blocks *u_pred, const uvg_b200_blocks *v_pred,
                                                const uvg_b200_blocks *u_rec, const uvg_b200_blocks *v_rec,
                                                int16_t *coeff_out, uint8_t *has_coeffs, size_t n_tus, void *stream)
{
  if (ensure_init()) return 1;
  if (n_tus == 0) return 0;
  if (!cfg || !u_ref || !v_ref || !u_pred || !v_pred || !u_rec || !v_rec || !coeff_out || !has_coeffs) {
    set_error("quant_cbcr_residual: NULL argument"); return 1;
  }
  if (joint_cb_cr < 1 || joint_cb_cr > 3) { set_error("quant_cbcr_residual: joint_cb_cr = %d", joint_cb_cr); return 1; }
  const int color = joint_cb_cr == 1 ? 2 : 1;                 // COLOR_V : COLOR_U
  return run_quantize_residual(cfg, cu_type, tr_idx, w, h, color, 0, early_skip, u_ref, u_pred, u_rec, v_ref, v_pred,
                               v_rec, joint_cb_cr * (jccr_sign ? -1 : 1), coeff_out, has_coeffs, nullptr, nullptr, 0,
                               n_tus, stream);
}

extern "C" int uvg_b200_coeff_abs_sum_dev(const int16_t *coeffs, size_t len, uint32_t *out, size_t n_tus,
                                          void *stream)
{
  if (ensure_init()) return 1;
  if (n_tus == 0) return 0;
  const size_t threads = n_tus * 32;
  coeff_sum_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      coeffs, len, out, n_tus, 0, 0, 0, 0, 0);
  UVGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int uvg_b200_fast_coeff_cost_dev(const int16_t *coeffs, int w, int h, uint64_t weights,
                                            uint32_t *out, size_t n_tus, void *stream)
{
  if (ensure_init()) return 1;
  if (n_tus == 0) return 0;
  const size_t threads = n_tus * 32;
  coeff_sum_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      coeffs, (size_t)w * h, out, n_tus, 1, (unsigned)(weights & 0xffff), (unsigned)((weights >> 16) & 0xffff),
      (unsigned)((weights >> 32) & 0xffff), (unsigned)((weights >> 48) & 0xffff));
  UVGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int uvg_b200_set_lfnst_tables(const int8_t *t8x8, const int8_t *t4x4)
{
  if (ensure_init()) return 1;
  if (!t8x8 || !t4x4) { set_error("set_lfnst_tables: NULL table"); return 1; }
  std::lock_guard<std::mutex> lk(lfnst_mu);
  if (!d_lfnst8) {
    int8_t *p8 = nullptr, *p4 = nullptr;
    UVGB_TRY(cudaMalloc(reinterpret_cast<void **>(&p8), 4 * 2 * 16 * 48));
    UVGB_TRY(cudaMalloc(reinterpret_cast<void **>(&p4), 4 * 2 * 16 * 16));
    d_lfnst8 = p8; d_lfnst4 = p4;
  }
  UVGB_TRY(cudaMemcpy(d_lfnst8, t8x8, 4 * 2 * 16 * 48, cudaMemcpyHostToDevice));
  UVGB_TRY(cudaMemcpy(d_lfnst4, t4x4, 4 * 2 * 16 * 16, cudaMemcpyHostToDevice));
  return 0;
}
