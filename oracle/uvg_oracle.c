/*
 * uvg_oracle.c - TEST INFRASTRUCTURE ONLY (see uvg_oracle.h).
 *
 * Plain-C restatement of the generic strategies of uvg266 for the distortion /
 * transform / quant hot path. Written from the behaviour of the reference, not from its
 * text: Hadamards are loops over butterfly stages, transforms are integer matrix
 * products with the reference's per-pass rounding (SURVEY.md Appendix D shows this is
 * bit-identical to the reference's partial-butterfly code), tables are generated from
 * the VVC base constants. Every function cites the reference lines it restates.
 */
#include "uvg_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define BD ORC_BIT_DEPTH
#define PIXEL_MAX ((1 << BD) - 1)

static inline int iabs(int x) { return x < 0 ? -x : x; }
static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

int orc_bit_depth(void) { return BD; }

/* ===================================================================================
 * strategies-picture
 * ================================================================================= */

/* reg_sad_generic, picture-generic.c:99-112. No bit-depth shift. */
unsigned orc_reg_sad(const orc_pixel *d1, const orc_pixel *d2, int w, int h,
                     unsigned stride1, unsigned stride2)
{
  unsigned sad = 0;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      sad += iabs((int)d1[y * stride1 + x] - (int)d2[y * stride2 + x]);
  return sad;
}

/* SAD_NXN, picture-generic.c:1054-1080: dense n*n, result >> (BD-8). */
unsigned orc_sad_nxn(int n, const orc_pixel *b1, const orc_pixel *b2)
{
  unsigned sum = 0;
  for (int i = 0; i < n * n; ++i) sum += iabs((int)b1[i] - (int)b2[i]);
  return sum >> (BD - 8);
}

/* SAD_DUAL_NXN, picture-generic.c:1091-1113. preds = pred_buffer: slot k at k*1024 px
 * (strategies-picture.h:48). */
void orc_sad_nxn_dual(int n, const orc_pixel *preds, const orc_pixel *orig, unsigned *costs_out)
{
  costs_out[0] = orc_sad_nxn(n, preds, orig);
  costs_out[1] = orc_sad_nxn(n, preds + 32 * 32, orig);
}

/* In-place Walsh-Hadamard transform of a w*h block of differences (w, h powers of two).
 * The reference's unrolled butterflies (hadamard_4x4_generic :118-200,
 * satd_8x8_subblock_generic :256-348, xCalcHADs16x8/8x16/4x8/8x4 :504-947) are full
 * Hadamards in various output orders; only sum|coef| and the DC term (all-plus row) are
 * used, and both are order independent. */
static void wht2d(int *d, int w, int h)
{
  for (int y = 0; y < h; ++y)
    for (int len = 1; len < w; len <<= 1)
      for (int i = 0; i < w; i += len << 1)
        for (int j = i; j < i + len; ++j) {
          int a = d[y * w + j], b = d[y * w + j + len];
          d[y * w + j] = a + b; d[y * w + j + len] = a - b;
        }
  for (int x = 0; x < w; ++x)
    for (int len = 1; len < h; len <<= 1)
      for (int i = 0; i < h; i += len << 1)
        for (int j = i; j < i + len; ++j) {
          int a = d[j * w + x], b = d[(j + len) * w + x];
          d[j * w + x] = a + b; d[(j + len) * w + x] = a - b;
        }
}

/* sum|coef| with the DC term replaced by |dc|>>2 (picture-generic.c:195-197, 342-343). */
static int had_abs_sum(const orc_pixel *o, int so, const orc_pixel *c, int sc, int w, int h)
{
  int d[16 * 8];
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      d[y * w + x] = (int)o[y * so + x] - (int)c[y * sc + x];
  wht2d(d, w, h);
  int s = 0;
  for (int i = 0; i < w * h; ++i) s += iabs(d[i]);
  s -= iabs(d[0]);
  s += iabs(d[0]) >> 2;
  return s;
}

/* uvg_satd_4x4_subblock_generic :217-230 -> hadamard_4x4_generic: (s+1)>>1 */
static unsigned satd4_sub(const orc_pixel *o, int so, const orc_pixel *c, int sc)
{
  return (unsigned)((had_abs_sum(o, so, c, sc, 4, 4) + 1) >> 1);
}
/* satd_8x8_subblock_generic :256-348: (s+2)>>2 */
static unsigned satd8_sub(const orc_pixel *o, int so, const orc_pixel *c, int sc)
{
  return (unsigned)((had_abs_sum(o, so, c, sc, 8, 8) + 2) >> 2);
}

/* satd_4x4_generic :205-212 (no BD shift). */
unsigned orc_satd_4x4(const orc_pixel *b1, const orc_pixel *b2)
{
  return satd4_sub(b1, 4, b2, 4);
}

/* SATD_NxN macro, strategies-picture.h:53-69 (n = 8..64; n = 4 -> satd_4x4). */
unsigned orc_satd_nxn(int n, const orc_pixel *b1, const orc_pixel *b2)
{
  if (n == 4) return orc_satd_4x4(b1, b2);
  unsigned sum = 0;
  for (int y = 0; y < n; y += 8)
    for (int x = 0; x < n; x += 8)
      sum += satd8_sub(&b1[y * n + x], n, &b2[y * n + x], n);
  return sum >> (BD - 8);
}

/* SATD_DUAL_NXN :371-410; the 4x4 dual has no BD shift (:401-405). */
void orc_satd_nxn_dual(int n, const orc_pixel *preds, const orc_pixel *orig, unsigned *costs_out)
{
  for (int k = 0; k < 2; ++k) {
    const orc_pixel *p = preds + k * 32 * 32;
    if (n == 4) { costs_out[k] = satd4_sub(orig, 4, p, 4); continue; }
    unsigned sum = 0;
    for (int y = 0; y < n; y += 8)
      for (int x = 0; x < n; x += 8)
        sum += satd8_sub(&orig[y * n + x], n, &p[y * n + x], n);
    costs_out[k] = sum >> (BD - 8);
  }
}

/* SATD_ANY_SIZE macro, strategies-picture.h:75-113. */
unsigned orc_satd_any_size(int w, int h, const orc_pixel *b1, int s1, const orc_pixel *b2, int s2)
{
  unsigned sum = 0;
  if (w % 8 != 0) {
    for (int y = 0; y < h; y += 4) sum += satd4_sub(&b1[y * s1], s1, &b2[y * s2], s2);
    b1 += 4; b2 += 4; w -= 4;
  }
  if (h % 8 != 0) {
    for (int x = 0; x < w; x += 4) sum += satd4_sub(&b1[x], s1, &b2[x], s2);
    b1 += 4 * s1; b2 += 4 * s2; h -= 4;
  }
  for (int y = 0; y < h; y += 8)
    for (int x = 0; x < w; x += 8)
      sum += satd8_sub(&b1[y * s1 + x], s1, &b2[y * s2 + x], s2);
  return sum >> (BD - 8);
}

/* SATD_ANY_SIZE_MULTI_GENERIC :412-479, including its literal edge handling: the 4x4 top
 * row starts at column 0 of the ORIGINAL pointers and the 8x8 loop restarts at
 * y = height%8 evaluated after height -= 4 (differs from the single version only when
 * both w%8 and h%8 are non-zero, or h%8 != 0 with h > 4). */
void orc_satd_any_size_quad(int w, int h, const orc_pixel *const preds[4], int stride,
                            const orc_pixel *orig, int orig_stride, unsigned *costs_out)
{
  unsigned acc[4] = { 0, 0, 0, 0 };
  const int wm8 = w % 8;
  if (wm8 != 0) {
    for (int y = 0; y < h; y += 4)
      for (int k = 0; k < 4; ++k)
        acc[k] += satd4_sub(&orig[y * orig_stride], orig_stride, &preds[k][y * stride], stride);
    w -= 4;
  }
  if (h % 8 != 0) {
    for (int x = 0; x < w; x += 4)
      for (int k = 0; k < 4; ++k)
        acc[k] += satd4_sub(&orig[x], orig_stride, &preds[k][x], stride);
    h -= 4;
  }
  for (int y = h % 8; y < h; y += 8)
    for (int x = wm8; x < w; x += 8)
      for (int k = 0; k < 4; ++k)
        acc[k] += satd8_sub(&orig[y * orig_stride + x], orig_stride,
                            &preds[k][y * stride + x], stride);
  for (int k = 0; k < 4; ++k) costs_out[k] = acc[k] >> (BD - 8);
}

/* xCalcHADs2x2 :481-501 (int16 intermediates; >>2 on the DC term only). */
static uint64_t had2x2(const orc_pixel *o, const orc_pixel *c, int so, int sc)
{
  int16_t d0 = (int16_t)(o[0] - c[0]), d1 = (int16_t)(o[1] - c[1]);
  int16_t d2 = (int16_t)(o[so] - c[sc]), d3 = (int16_t)(o[so + 1] - c[sc + 1]);
  int16_t m0 = (int16_t)(d0 + d2), m1 = (int16_t)(d1 + d3);
  int16_t m2 = (int16_t)(d0 - d2), m3 = (int16_t)(d1 - d3);
  uint64_t s = 0;
  s += (uint64_t)(iabs(m0 + m1) >> 2);
  s += (uint64_t)iabs(m0 - m1);
  s += (uint64_t)iabs(m2 + m3);
  s += (uint64_t)iabs(m2 - m3);
  return s;
}

/* Rectangular kernels: sad = (int)(sad / sqrt(W*H_kernel) * 2) in IEEE double
 * (:647, :786 sqrt(16.0*8); :862, :943 sqrt(4.0*8)). */
static uint64_t had_rect(const orc_pixel *o, int so, const orc_pixel *c, int sc, int w, int h)
{
  int s = had_abs_sum(o, so, c, sc, w, h);
  double norm = (w * h == 128) ? sqrt(16.0 * 8) : sqrt(4.0 * 8);
  s = (int)(s / norm * 2);
  return (uint64_t)(int64_t)s;
}

/* xGetHADs :949-1049 (= satd_any_size_vtm). Returns uint64 truncated to unsigned by the
 * function-pointer type; no BD shift. */
unsigned orc_satd_any_size_vtm(int w, int h, const orc_pixel *org, int so,
                               const orc_pixel *cur, int sc)
{
  uint64_t sum = 0;
  int bw, bh, kind;
  if (w > h && (h & 7) == 0 && (w & 15) == 0) { bw = 16; bh = 8; kind = 0; }
  else if (w < h && (w & 7) == 0 && (h & 15) == 0) { bw = 8; bh = 16; kind = 0; }
  else if (w > h && (h & 3) == 0 && (w & 7) == 0) { bw = 8; bh = 4; kind = 0; }
  else if (w < h && (w & 3) == 0 && (h & 7) == 0) { bw = 4; bh = 8; kind = 0; }
  else if (h % 8 == 0 && w % 8 == 0) { bw = 8; bh = 8; kind = 1; }
  else if (h % 4 == 0 && w % 4 == 0) { bw = 4; bh = 4; kind = 2; }
  else if (h % 2 == 0 && w % 2 == 0) { bw = 2; bh = 2; kind = 3; }
  else return 0;
  for (int y = 0; y < h; y += bh)
    for (int x = 0; x < w; x += bw) {
      const orc_pixel *o = &org[y * so + x], *c = &cur[y * sc + x];
      switch (kind) {
        case 0: sum += had_rect(o, so, c, sc, bw, bh); break;
        case 1: sum += satd8_sub(o, so, c, sc); break;
        case 2: sum += satd4_sub(o, so, c, sc); break;
        default: sum += had2x2(o, c, so, sc); break;
      }
    }
  return (unsigned)sum;
}

/* pixels_calc_ssd_generic :1115-1130: accumulates in (wrapping) int, >> 2*(BD-8). */
unsigned orc_pixels_calc_ssd(const orc_pixel *ref, const orc_pixel *rec, int ref_stride,
                             int rec_stride, int w, int h)
{
  uint32_t ssd = 0; /* unsigned accumulate == two's complement wrap of the int */
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int diff = (int)ref[x + y * ref_stride] - (int)rec[x + y * rec_stride];
      ssd += (uint32_t)(diff * diff);
    }
  return (unsigned)((int32_t)ssd >> (2 * (BD - 8)));
}

/* ver_sad_generic :1266-1279 */
uint32_t orc_ver_sad(const orc_pixel *pic, const orc_pixel *ref_row, int w, int h,
                     uint32_t pic_stride)
{
  uint32_t sad = 0;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      sad += iabs((int)pic[y * pic_stride + x] - (int)ref_row[x]);
  return sad;
}

static uint32_t hor_sad_col(const orc_pixel *pic, const orc_pixel *ref, int w, int h,
                            uint32_t ps, uint32_t rs)
{
  uint32_t sad = 0; /* static hor_sad :1292-1306: one reference column replicated */
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      sad += iabs((int)pic[y * ps + x] - (int)ref[y * rs]);
  return sad;
}

/* hor_sad_generic :1309-1331 */
uint32_t orc_hor_sad(const orc_pixel *pic, const orc_pixel *ref, int w, int h,
                     uint32_t ps, uint32_t rs, uint32_t left, uint32_t right)
{
  uint32_t r = 0;
  if (left) {
    r += hor_sad_col(pic, ref + left, (int)left, h, ps, rs);
    r += orc_reg_sad(pic + left, ref + left, w - (int)left, h, ps, rs);
  } else if (right) {
    r += orc_reg_sad(pic, ref, w - (int)right, h, ps, rs);
    r += hor_sad_col(pic + w - right, ref + w - right - 1, (int)right, h, ps, rs);
  } else {
    r += orc_reg_sad(pic, ref, w, h, ps, rs);
  }
  return r;
}

/* generate_residual_generic :1360-1369: dense output, stride = width. */
void orc_generate_residual(const orc_pixel *ref, const orc_pixel *pred, int16_t *residual,
                           int w, int h, int ref_stride, int pred_stride)
{
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      residual[x + y * w] = (int16_t)((int)ref[x + y * ref_stride] - (int)pred[x + y * pred_stride]);
}

/* uvg_image_calc_sad, image.c:438-473 with image_interpolated_sad :310-427: blocks that
 * leave the reference frame are compared against the nearest edge pixel (cor_sad /
 * ver_sad / hor_sad / reg_sad split) == SAD against the edge-replicated reference.
 * Result >> (BD-8) (:472). Pinned by tests/sad_tests.c:144-282. */
unsigned orc_image_calc_sad(const orc_pixel *pic, const orc_pixel *ref, int pic_w, int pic_h,
                            int stride, int pic_x, int pic_y, int ref_x, int ref_y,
                            int bw, int bh)
{
  unsigned sad = 0;
  for (int y = 0; y < bh; ++y)
    for (int x = 0; x < bw; ++x) {
      int rx = clip3(0, pic_w - 1, ref_x + x), ry = clip3(0, pic_h - 1, ref_y + y);
      sad += iabs((int)pic[(pic_y + y) * stride + pic_x + x] - (int)ref[ry * stride + rx]);
    }
  return sad >> (BD - 8);
}

/* ===================================================================================
 * strategies-dct
 * ================================================================================= */

/* DCT-2 base magnitudes round(64*sqrt(2)*cos(m*pi/64)) as tuned by HEVC/VVC, m = 0..31
 * (m = 0 entry is the 64 of the DC row); same numbers as uvg_g_dct_32 / DEFINE_DCT2_P32
 * (dct-generic.c:52-126, :774-855). */
static const int8_t k_dct2_mag[32] = { 64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73,
  70, 67, 64, 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4 };
/* DST-7 magnitudes c_j ~ sin(j*pi/(2N+1)), j = 1..N (VVC tuned; dct-generic.c:1018-1027). */
static const int8_t k_dst7_mag4[4] = { 29, 55, 74, 84 };
static const int8_t k_dst7_mag8[8] = { 17, 32, 46, 60, 71, 78, 85, 86 };
static const int8_t k_dst7_mag16[16] = { 8, 17, 25, 33, 40, 48, 55, 62, 68, 73, 77, 81, 85, 87, 88, 88 };
static const int8_t k_dst7_mag32[32] = { 4, 9, 13, 17, 21, 26, 30, 34, 38, 42, 46, 50, 53, 56,
  60, 63, 66, 68, 72, 74, 77, 78, 80, 82, 84, 85, 86, 87, 88, 89, 90, 90 };

static int dct2_entry(int n, int k, int i)
{
  if (k == 0) return 64;
  int m = ((2 * i + 1) * k * (32 / n)) % 128; /* angle m*pi/64 */
  if (m > 64) m = 128 - m;
  if (m == 32) return 0;
  return m < 32 ? k_dct2_mag[m] : -k_dct2_mag[64 - m];
}

static int dst7_entry(int n, int k, int i)
{
  const int8_t *mag = n == 4 ? k_dst7_mag4 : n == 8 ? k_dst7_mag8 : n == 16 ? k_dst7_mag16 : k_dst7_mag32;
  const int p = 2 * n + 1;
  int m = ((2 * k + 1) * (i + 1)) % (2 * p); /* sin(m*pi/p) */
  int sign = 1;
  if (m > p) { m -= p; sign = -1; }
  if (m > n) m = p - m;
  if (m == 0) return 0;
  return sign * mag[m - 1];
}

int orc_tr_matrix(int type, int n, int16_t *out)
{
  if (n != 2 && n != 4 && n != 8 && n != 16 && n != 32) return 0;
  if (n == 2 && type != ORC_DCT2) return 0;
  for (int k = 0; k < n; ++k)
    for (int i = 0; i < n; ++i) {
      int v;
      if (type == ORC_DCT2) v = dct2_entry(n, k, i);
      else if (type == ORC_DST7) v = dst7_entry(n, k, i);
      else v = ((k & 1) ? -1 : 1) * dst7_entry(n, k, n - 1 - i); /* DCT-8 = flipped DST-7 */
      out[k * n + i] = (int16_t)v;
    }
  return 1;
}

static inline int rnd_shift(int x, int shift)
{
  int add = shift > 0 ? 1 << (shift - 1) : 0;
  return (x + add) >> shift;
}

/* One forward 1-D pass in the reference's kernel contract
 * (src, dst, shift, line, skip_line, skip_line2), e.g. fastForwardDCT2_B8 :1159,
 * fastForwardDST7_B16 :1574: transform `line - skip_line` rows of N samples, write
 * transposed (dst[k*line + j]), plain int16 conversion (wraps); rows beyond reduced_line
 * and - for the 8/16/32-point DST7/DCT8 kernels only (:1574-2343; the 4-point and all DCT-2
 * kernels ignore skip_line2) - the last skip_line2 frequencies are zero. */
static void fwd_pass(int type, int n, const int16_t *src, int16_t *dst, int shift, int line,
                     int skip_line, int skip_line2)
{
  int16_t T[32 * 32];
  orc_tr_matrix(type, n, T);
  const int reduced = line - skip_line;
  const int cutoff = (type == ORC_DCT2 || n < 8) ? n : n - skip_line2;
  memset(dst, 0, sizeof(int16_t) * (size_t)n * line);
  for (int j = 0; j < reduced; ++j)
    for (int k = 0; k < cutoff; ++k) {
      int sum = 0;
      for (int i = 0; i < n; ++i) sum += (int)src[j * n + i] * T[k * n + i];
      dst[k * line + j] = (int16_t)rnd_shift(sum, shift);
    }
}

/* One inverse 1-D pass (fastInverseDCT2_B8 :1203, fastInverseDST7_B16 :1632 ...):
 * dst[j*N + i] = clip16(rnd(sum_k src[k*line + j] * T[k*N + i])), rows >= reduced zero.
 * Only the 8-point DST7/DCT8 inverse kernels (:2272, :2343) stop the sum at
 * N - skip_line2; every other inverse kernel reads all N inputs. */
static void inv_pass(int type, int n, const int16_t *src, int16_t *dst, int shift, int line,
                     int skip_line, int skip_line2)
{
  int16_t T[32 * 32];
  orc_tr_matrix(type, n, T);
  const int reduced = line - skip_line;
  const int cutoff = (type != ORC_DCT2 && n == 8) ? n - skip_line2 : n;
  memset(dst, 0, sizeof(int16_t) * (size_t)n * line);
  for (int j = 0; j < reduced; ++j)
    for (int i = 0; i < n; ++i) {
      int sum = 0;
      for (int k = 0; k < cutoff; ++k) sum += (int)src[k * line + j] * T[k * n + i];
      dst[j * n + i] = (int16_t)clip3(-32768, 32767, rnd_shift(sum, shift));
    }
}

/* DCT_NXN_GENERIC :720-729 (partial_butterfly_N_generic twice). */
void orc_dct_nxn(int n, int bitdepth, const int16_t *in, int16_t *out)
{
  int16_t tmp[32 * 32];
  const int l2 = ilog2(n);
  fwd_pass(ORC_DCT2, n, in, tmp, l2 - 1 + (bitdepth - 8), n, 0, 0);
  fwd_pass(ORC_DCT2, n, tmp, out, l2 + 6, n, 0, 0);
}

/* IDCT_NXN_GENERIC :731-740. */
void orc_idct_nxn(int n, int bitdepth, const int16_t *in, int16_t *out)
{
  int16_t tmp[32 * 32];
  inv_pass(ORC_DCT2, n, in, tmp, 7, n, 0, 0);
  inv_pass(ORC_DCT2, n, tmp, out, 12 - (bitdepth - 8), n, 0, 0);
}

/* fast_forward_dst_4x4_generic :752-760 -> fast_forward_dst_4_generic :359-375. */
void orc_fast_forward_dst_4x4(int bitdepth, const int16_t *in, int16_t *out)
{
  int16_t tmp[16];
  fwd_pass(ORC_DST7, 4, in, tmp, 1 + (bitdepth - 8), 4, 0, 0);
  fwd_pass(ORC_DST7, 4, tmp, out, 8, 4, 0, 0);
}

/* fast_inverse_dst_4x4_generic :762-770 -> fast_inverse_dst_4_generic :377-393. */
void orc_fast_inverse_dst_4x4(int bitdepth, const int16_t *in, int16_t *out)
{
  int16_t tmp[16];
  inv_pass(ORC_DST7, 4, in, tmp, 7, 4, 0, 0);
  inv_pass(ORC_DST7, 4, tmp, out, 12 - (bitdepth - 8), 4, 0, 0);
}

/* uvg_get_tr_type, dct-generic.c:2501-2557. */
void orc_get_tr_type(int w, int h, int color, int cu_type, int tr_idx, int lfnst_idx,
                     int cr_lfnst_idx, int isp_mode, int mts_type, int *hor, int *ver)
{
  *hor = ORC_DCT2; *ver = ORC_DCT2;
  if (color != 0) return;
  const int intra = cu_type == ORC_CU_INTRA;
  const int explicit_mts = mts_type == ORC_MTS_BOTH ||
      (intra ? mts_type == ORC_MTS_INTRA : (mts_type == ORC_MTS_INTER && cu_type == ORC_CU_INTER));
  const int implicit_mts = intra && (mts_type == ORC_MTS_IMPLICIT || mts_type == ORC_MTS_INTER);
  const int is_isp = (intra && isp_mode) ? 1 : 0;
  const int lfnst = color == 0 ? lfnst_idx : cr_lfnst_idx;
  if (is_isp && lfnst) return;
  if (implicit_mts || (is_isp && explicit_mts)) {
    if (w >= 4 && w <= 16) *hor = ORC_DST7;
    if (h >= 4 && h <= 16) *ver = ORC_DST7;
    return;
  }
  if (explicit_mts && tr_idx > 1 /* MTS_SKIP */) {
    static const int subset[4][2] = { { ORC_DST7, ORC_DST7 }, { ORC_DCT8, ORC_DST7 },
                                      { ORC_DST7, ORC_DCT8 }, { ORC_DCT8, ORC_DCT8 } };
    *hor = subset[tr_idx - 2][0];
    *ver = subset[tr_idx - 2][1];
  }
}

static void mts_skips(int w, int h, int color, int lfnst_idx, int cr_lfnst_idx, int th, int tv,
                      int *skip_w, int *skip_h)
{
  *skip_w = (th != ORC_DCT2 && w == 32) ? 16 : (w > 32 ? w - 32 : 0);
  *skip_h = (tv != ORC_DCT2 && h == 32) ? 16 : (h > 32 ? h - 32 : 0);
  if ((lfnst_idx && color == 0) || (cr_lfnst_idx && color != 0)) {
    if ((w == 4 && h > 4) || (w > 4 && h == 4)) { *skip_w = w - 4; *skip_h = h - 4; }
    else if (w >= 8 && h >= 8) { *skip_w = w - 8; *skip_h = h - 8; }
  }
}

/* mts_dct_generic, dct-generic.c:2560-2618. */
void orc_mts_dct(int bitdepth, int color, int cu_type, int tr_idx, int lfnst_idx,
                 int cr_lfnst_idx, int isp_mode, int w, int h, const int16_t *in,
                 int16_t *out, int mts_type)
{
  int th, tv;
  orc_get_tr_type(w, h, color, cu_type, tr_idx, lfnst_idx, cr_lfnst_idx, isp_mode, mts_type, &th, &tv);
  if (th == ORC_DCT2 && tv == ORC_DCT2 && !lfnst_idx && !cr_lfnst_idx && w == h) {
    orc_dct_nxn(w, bitdepth, in, out);
    return;
  }
  int skip_w, skip_h;
  mts_skips(w, h, color, lfnst_idx, cr_lfnst_idx, th, tv, &skip_w, &skip_h);
  const int l2w1 = ilog2(w) - 1, l2h1 = ilog2(h) - 1;
  int16_t tmp[32 * 32];
  if (h == 1) {
    fwd_pass(th, w, in, out, l2w1 + bitdepth - 8, 1, 0, skip_w);
  } else if (w == 1) {
    fwd_pass(tv, h, in, out, l2h1 + 1 + bitdepth + 6 - 15, 1, 0, skip_h);
  } else {
    fwd_pass(th, w, in, tmp, l2w1 + bitdepth - 8, h, 0, skip_w);
    fwd_pass(tv, h, tmp, out, l2h1 + 7, w, skip_w, skip_h);
  }
}

/* mts_idct_generic, dct-generic.c:2621-2678. */
void orc_mts_idct(int bitdepth, int color, int cu_type, int tr_idx, int lfnst_idx,
                  int cr_lfnst_idx, int isp_mode, int w, int h, const int16_t *in,
                  int16_t *out, int mts_type)
{
  int th, tv;
  orc_get_tr_type(w, h, color, cu_type, tr_idx, lfnst_idx, cr_lfnst_idx, isp_mode, mts_type, &th, &tv);
  if (th == ORC_DCT2 && tv == ORC_DCT2 && !lfnst_idx && !cr_lfnst_idx && w == h) {
    orc_idct_nxn(w, bitdepth, in, out);
    return;
  }
  int skip_w, skip_h;
  mts_skips(w, h, color, lfnst_idx, cr_lfnst_idx, th, tv, &skip_w, &skip_h);
  const int shift1 = 7, shift2 = 20 - bitdepth;
  int16_t tmp[32 * 32];
  if (h == 1) {
    inv_pass(th, w, in, out, shift2 + 1, 1, 0, skip_w);
  } else if (w == 1) {
    inv_pass(tv, h, in, out, shift2 + 1, 1, 0, skip_h);
  } else {
    inv_pass(tv, h, in, tmp, shift1, w, skip_w, skip_h);
    inv_pass(th, w, tmp, out, shift2, h, 0, skip_w);
  }
}

/* ===================================================================================
 * strategies-quant
 * ================================================================================= */

/* uvg_g_log2_sbb_size (tables.c:13-24): coefficient-group log2 dims per block log2 dims. */
static void cg_log2(int l2w, int l2h, int *cw, int *ch)
{
  if (l2w >= 2 && l2h >= 2) { *cw = 2; *ch = 2; return; }
  /* one dimension < 4: a 16-coefficient (or smaller) group stretched along the other */
  if (l2w == 0) { *cw = 0; *ch = l2h < 4 ? l2h : 4; return; }
  if (l2h == 0) { *ch = 0; *cw = l2w < 4 ? l2w : 4; return; }
  if (l2w == 1) { *cw = 1; *ch = (l2h == 2) ? 1 : (l2h >= 3 ? 3 : l2h); return; }
  /* l2h == 1 */
  *ch = 1; *cw = (l2w == 2) ? 1 : (l2w >= 3 ? 3 : l2w);
}

static int diag_fill(int w, int h, int *xs, int *ys)
{
  int n = 0; /* up-right diagonal scan: each anti-diagonal from bottom-left to top-right */
  for (int s = 0; s < w + h - 1; ++s)
    for (int x = 0; x < w; ++x) {
      int y = s - x;
      if (y >= 0 && y < h) { xs[n] = x; ys[n] = y; ++n; }
    }
  return n;
}

/* g_scan_order[SCAN_GROUP_4X4][log2w][log2h] (tables.c, via uvg_get_scan_order_table
 * :2646): coefficient groups in diagonal order, diagonal order inside each group. */
void orc_scan_table(int l2w, int l2h, uint32_t *out)
{
  int cw, ch;
  cg_log2(l2w, l2h, &cw, &ch);
  const int w = 1 << l2w, gw = 1 << cw, gh = 1 << ch;
  static int gx[1024], gy[1024], ix[1024], iy[1024];
  int ng = diag_fill(w >> cw, (1 << l2h) >> ch, gx, gy);
  int ni = diag_fill(gw, gh, ix, iy);
  int n = 0;
  for (int g = 0; g < ng; ++g)
    for (int i = 0; i < ni; ++i)
      out[n++] = (uint32_t)((gy[g] * gh + iy[i]) * w + gx[g] * gw + ix[i]);
}

/* uvg_get_scaled_qp, transform.c:150-165. */
int32_t orc_get_scaled_qp(int color, int qp, int qp_offset, const int8_t *chroma_map)
{
  if (color == 0) return qp + qp_offset;
  int q = clip3(-qp_offset, 57, qp);
  return chroma_map ? chroma_map[qp] + qp_offset : q + qp_offset;
}

static const int16_t k_quant_scales[2][6] = { { 26214, 23302, 20560, 18396, 16384, 14564 },
                                              { 18396, 16384, 14564, 13107, 11651, 10280 } };
static const int16_t k_inv_quant_scales[2][6] = { { 40, 45, 51, 57, 64, 72 },
                                                  { 57, 64, 72, 80, 90, 102 } };

/* uvg_quant_generic, quant-generic.c:51-234, scaling lists off (flat table =
 * uvg_g_quant_scales[0][qp%6], scalinglist.c:416). */
void orc_quant(const orc_quant_params *p, const int16_t *coef, int16_t *q_coef, int w, int h,
               int color, int block_type, int transform_skip, int lfnst_idx)
{
  (void)block_type;
  const int l2w = ilog2(w), l2h = ilog2(h);
  uint32_t scan[32 * 32];
  orc_scan_table(l2w, l2h, scan);
  int32_t qp_scaled = orc_get_scaled_qp(color, p->qp, (p->bitdepth - 8) * 6, p->chroma_qp_map);
  if (transform_skip && qp_scaled < 4 + 6 * 2) qp_scaled = 4 + 6 * 2; /* MIN_QP_PRIME_TS=2 */
  const int needs_scale = !transform_skip && ((l2w + l2h) % 2 == 1);
  const int32_t flat_list_coeff = k_quant_scales[0][qp_scaled % 6]; /* quant_coeff[] when lists are off */
  const int32_t transform_shift = 15 - p->bitdepth - ((l2w + l2h) >> 1) - needs_scale;
  const int64_t q_bits = 14 + qp_scaled / 6 + (transform_skip ? 0 : transform_shift);
  const int32_t add = (p->slice_intra ? 171 : 85) << (q_bits - 9);
  const int32_t q_bits8 = (int32_t)q_bits - 8;
  const int32_t default_coeff = k_quant_scales[needs_scale][qp_scaled % 6];
  uint32_t ac_sum = 0;
  const int max_lfnst = ((w == 4 && h == 4) || (w == 8 && h == 8)) ? 8 : 16;

  if (lfnst_idx == 0) {
    for (int n = 0; n < w * h; ++n) {
      int32_t level = coef[n];
      int64_t abs_level = level < 0 ? -(int64_t)level : level;
      int32_t sign = level < 0 ? -1 : 1;
      level = (int32_t)((abs_level * default_coeff + add) >> q_bits);
      ac_sum += (uint32_t)level;
      level *= sign;
      q_coef[n] = (int16_t)clip3(-32768, 32767, level);
    }
  } else {
    memset(q_coef, 0, sizeof(int16_t) * (size_t)w * h);
    for (int n = 0; n < max_lfnst; ++n) {
      const uint32_t idx = scan[n];
      int32_t level = coef[idx];
      int64_t abs_level = level < 0 ? -(int64_t)level : level;
      int32_t sign = level < 0 ? -1 : 1;
      level = (int32_t)((abs_level * flat_list_coeff + add) >> q_bits); /* always the list, :114 */
      ac_sum += (uint32_t)level;
      level *= sign;
      q_coef[idx] = (int16_t)clip3(-32768, 32767, level);
    }
  }

  if (!p->signhide || ac_sum < 2) return;

  /* Sign-bit hiding :124-233. delta_u always uses the list table (flat row 0). */
  int32_t delta_u[32 * 32];
  memset(delta_u, 0, sizeof(delta_u));
  if (lfnst_idx == 0) {
    for (int n = 0; n < w * h; ++n) {
      int32_t c = coef[n];
      int64_t abs_level = c < 0 ? -(int64_t)c : c;
      int32_t level = (int32_t)((abs_level * flat_list_coeff + add) >> q_bits);
      delta_u[n] = (int32_t)((abs_level * flat_list_coeff - ((int64_t)level << q_bits)) >> q_bits8);
    }
  } else {
    for (int n = 0; n < max_lfnst; ++n) {
      const uint32_t idx = scan[n];
      int32_t c = coef[idx];
      int64_t abs_level = c < 0 ? -(int64_t)c : c;
      int32_t level = (int32_t)((abs_level * flat_list_coeff + add) >> q_bits);
      delta_u[idx] = (int32_t)((abs_level * flat_list_coeff - ((int64_t)level << q_bits)) >> q_bits8);
    }
  }

  int last_cg = -1;
  for (int subset = (w * h - 1) >> 4; subset >= 0; --subset) {
    const int subpos = subset << 4;
    int first_nz = 16, last_nz = -1, abssum = 0;
    for (int n = 15; n >= 0; --n) if (q_coef[scan[n + subpos]]) { last_nz = n; break; }
    for (int n = 0; n < 16; ++n) if (q_coef[scan[n + subpos]]) { first_nz = n; break; }
    for (int n = first_nz; n <= last_nz; ++n) abssum += q_coef[scan[n + subpos]];
    if (last_nz >= 0 && last_cg == -1) last_cg = 1;
    if (last_nz - first_nz >= 4) {
      const int signbit = q_coef[scan[subpos + first_nz]] > 0 ? 0 : 1;
      if (signbit != (abssum & 1)) {
        int32_t min_cost = 0x7fffffff, cur_cost = 0x7fffffff;
        int min_pos = -1;
        int16_t final_change = 0, cur_change = 0;
        for (int n = (last_cg == 1 ? last_nz : 15); n >= 0; --n) {
          const uint32_t pos = scan[n + subpos];
          if (q_coef[pos] != 0) {
            if (delta_u[pos] > 0) { cur_cost = -delta_u[pos]; cur_change = 1; }
            else if (n == first_nz && iabs(q_coef[pos]) == 1) { cur_cost = 0x7fffffff; }
            else { cur_cost = delta_u[pos]; cur_change = -1; }
          } else if (n < first_nz && ((coef[pos] >= 0) ? 0 : 1) != signbit) {
            cur_cost = 0x7fffffff;
          } else {
            cur_cost = -delta_u[pos]; cur_change = 1;
          }
          if (cur_cost < min_cost) { min_cost = cur_cost; final_change = cur_change; min_pos = (int)pos; }
        }
        if (q_coef[min_pos] == 32767 || q_coef[min_pos] == -32768) final_change = -1;
        if (coef[min_pos] >= 0) q_coef[min_pos] = (int16_t)(q_coef[min_pos] + final_change);
        else q_coef[min_pos] = (int16_t)(q_coef[min_pos] - final_change);
      }
    }
    if (last_cg == 1) last_cg = 0;
  }
}

/* uvg_dequant_generic, quant-generic.c:618-669, flat branch (:660-668). */
void orc_dequant(const orc_quant_params *p, const int16_t *q_coef, int16_t *coef, int w, int h,
                 int color, int block_type, int transform_skip)
{
  (void)block_type;
  const int l2w = ilog2(w), l2h = ilog2(h);
  const int32_t transform_shift = 15 - p->bitdepth - ((l2w + l2h) >> 1);
  const int needs_scale = !transform_skip && ((l2w + l2h) % 2 == 1);
  int32_t qp_scaled = orc_get_scaled_qp(color, p->qp, (p->bitdepth - 8) * 6, p->chroma_qp_map);
  if (transform_skip && qp_scaled < 4 + 6 * 2) qp_scaled = 4 + 6 * 2;
  const int32_t shift = 20 - 14 - (transform_skip ? 0 : transform_shift - needs_scale);
  const int32_t scale = k_inv_quant_scales[needs_scale][qp_scaled % 6] << (qp_scaled / 6);
  const int32_t add = 1 << (shift - 1);
  for (int n = 0; n < w * h; ++n) {
    int32_t c = ((int32_t)q_coef[n] * scale + add) >> shift;
    coef[n] = (int16_t)clip3(-32768, 32767, c);
  }
}

/* uvg_quantize_residual_generic, quant-generic.c:460-612, for the plain path
 * (no LMCS chroma adj, no LFNST, quant = uvg_quant). transform2d/itransform2d:
 * transform.c:256-292. */
int orc_quantize_residual(const orc_quant_params *p, int cu_type, int tr_idx, int w, int h,
                          int color, int use_trskip, int in_stride, int out_stride,
                          const orc_pixel *ref_in, const orc_pixel *pred_in,
                          orc_pixel *rec_out, int16_t *coeff_out, int early_skip)
{
  int16_t residual[32 * 32], coeff[32 * 32];
  orc_generate_residual(ref_in, pred_in, residual, w, h, in_stride, in_stride);
  if (use_trskip) {
    memcpy(coeff, residual, sizeof(int16_t) * (size_t)w * h); /* uvg_transformskip :217 */
  } else if (p->mts || w != h) {
    orc_mts_dct(p->bitdepth, color, cu_type, tr_idx, 0, 0, p->isp_mode, w, h, residual, coeff, p->mts);
  } else {
    orc_dct_nxn(w, p->bitdepth, residual, coeff);
  }
  const int ts_flag = (tr_idx == 1 /* MTS_SKIP */ && color == 0);
  orc_quant(p, coeff, coeff_out, w, h, color, cu_type, ts_flag, 0);
  int has_coeffs = 0;
  for (int i = 0; i < w * h; ++i) if (coeff_out[i] != 0) { has_coeffs = 1; break; }
  if (has_coeffs && !early_skip) {
    orc_dequant(p, coeff_out, coeff, w, h, color, cu_type, ts_flag);
    if (use_trskip) {
      memcpy(residual, coeff, sizeof(int16_t) * (size_t)w * h);
    } else if (p->mts || w != h) {
      orc_mts_idct(p->bitdepth, color, cu_type, tr_idx, 0, 0, p->isp_mode, w, h, coeff, residual, p->mts);
    } else {
      orc_idct_nxn(w, p->bitdepth, coeff, residual);
    }
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        int16_t val = (int16_t)(residual[x + y * w] + pred_in[x + y * in_stride]);
        rec_out[x + y * out_stride] = (orc_pixel)clip3(0, PIXEL_MAX, val);
      }
  } else if (rec_out != pred_in) {
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x)
        rec_out[x + y * out_stride] = pred_in[x + y * in_stride];
  }
  return has_coeffs;
}

/* uvg_quant_cbcr_residual_generic, quant-generic.c:241-442, plain path (no LFNST, quant = uvg_quant,
 * no LMCS). cbf_mask = joint_cb_cr * (jccr_sign ? -1 : 1) (:273); the coded plane is V for joint_cb_cr == 1,
 * else U (:311, :337). Both reconstructions are always written (:415-441). Returns joint_cb_cr or 0 (:444). */
int orc_quant_cbcr_residual(const orc_quant_params *p, int cu_type, int tr_idx, int w, int h, int joint_cb_cr,
                            int jccr_sign, int in_stride, int out_stride, const orc_pixel *u_ref,
                            const orc_pixel *v_ref, const orc_pixel *u_pred, const orc_pixel *v_pred,
                            orc_pixel *u_rec, orc_pixel *v_rec, int16_t *coeff_out, int early_skip)
{
  int16_t u_res[32 * 32], v_res[32 * 32], comb[32 * 32], coeff[32 * 32];
  orc_generate_residual(u_ref, u_pred, u_res, w, h, in_stride, in_stride);
  orc_generate_residual(v_ref, v_pred, v_res, w, h, in_stride, in_stride);
  const int mask = joint_cb_cr * (jccr_sign ? -1 : 1);
  for (int i = 0; i < w * h; ++i) {
    const int cb = u_res[i], cr = v_res[i];
    int v;
    switch (mask) {                                   /* :275-309, C division truncates toward zero */
      case 2: v = (4 * cb + 2 * cr) / 5; break;
      case -2: v = (4 * cb - 2 * cr) / 5; break;
      case 3: v = (cb + cr) / 2; break;
      case -3: v = (cb - cr) / 2; break;
      case 1: v = (4 * cr + 2 * cb) / 5; break;
      default: v = (4 * cr - 2 * cb) / 5; break;
    }
    comb[i] = (int16_t)v;
  }
  const int color = joint_cb_cr == 1 ? 2 : 1;
  if (p->mts || w != h) orc_mts_dct(p->bitdepth, color, cu_type, tr_idx, 0, 0, 0, w, h, comb, coeff, p->mts);
  else orc_dct_nxn(w, p->bitdepth, comb, coeff);
  orc_quant(p, coeff, coeff_out, w, h, color, cu_type, 0, 0);
  int has_coeffs = 0;
  for (int i = 0; i < w * h; ++i) if (coeff_out[i] != 0) { has_coeffs = 1; break; }
  if (has_coeffs && !early_skip) {
    orc_dequant(p, coeff_out, coeff, w, h, color, cu_type, 0);
    if (p->mts || w != h) orc_mts_idct(p->bitdepth, color, cu_type, tr_idx, 0, 0, 0, w, h, coeff, comb, p->mts);
    else orc_idct_nxn(w, p->bitdepth, coeff, comb);
    for (int i = 0; i < w * h; ++i) {                 /* :388-414 */
      const int c = comb[i];
      switch (mask) {
        case 2: u_res[i] = (int16_t)c; v_res[i] = (int16_t)(c >> 1); break;
        case -2: u_res[i] = (int16_t)c; v_res[i] = (int16_t)(-c >> 1); break;
        case 3: u_res[i] = (int16_t)c; v_res[i] = (int16_t)c; break;
        case -3: u_res[i] = (int16_t)c; v_res[i] = (int16_t)(-c); break;
        case 1: u_res[i] = (int16_t)(c >> 1); v_res[i] = (int16_t)c; break;
        default: u_res[i] = (int16_t)(-c >> 1); v_res[i] = (int16_t)c; break;
      }
    }
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        const int16_t uv = (int16_t)(u_res[x + y * w] + u_pred[x + y * in_stride]);
        const int16_t vv = (int16_t)(v_res[x + y * w] + v_pred[x + y * in_stride]);
        u_rec[x + y * out_stride] = (orc_pixel)clip3(0, PIXEL_MAX, uv);
        v_rec[x + y * out_stride] = (orc_pixel)clip3(0, PIXEL_MAX, vv);
      }
  } else {
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        u_rec[x + y * out_stride] = u_pred[x + y * in_stride];
        v_rec[x + y * out_stride] = v_pred[x + y * in_stride];
      }
  }
  return has_coeffs ? joint_cb_cr : 0;
}

/* ===================================================================================
 * LFNST (first piece): uvg_fwd_lfnst / uvg_inv_lfnst, transform.c:878-1200.
 * The two trained kernels are NOT part of this repository: the test harness takes their
 * addresses from the compiled reference (exported read-only symbols uvg_lfnst_8x8 /
 * uvg_lfnst_4x4, lfnst_tables.h:56/:211) and hands them in through orc_set_lfnst_tables.
 * Everything else (set selection, transposition, scan positions) follows from the code cited.
 * ================================================================================= */
static const int8_t *g_lfnst8;   /* [4][2][16][48] */
static const int8_t *g_lfnst4;   /* [4][2][16][16] */
void orc_set_lfnst_tables(const int8_t *t8x8, const int8_t *t4x4) { g_lfnst8 = t8x8; g_lfnst4 = t4x4; }

/* uvg_lfnst_lut (lfnst_tables.h:51-54) as intervals of the extended mode index 0..94 */
static int lfnst_set(int m) { return m < 2 ? 0 : m <= 12 ? 1 : m <= 23 ? 2 : m <= 44 ? 3 : m <= 55 ? 2 : 1; }

/* uvg_wide_angle_correction, intra.c:637-658 */
static int wide_angle(int mode, int l2w, int l2h, int account_for_dc_planar)
{
  static const int shift[6] = { 0, 6, 10, 12, 14, 15 };
  if (l2w != l2h && mode > 1 && mode <= 66) {
    const int d = l2w > l2h ? l2w - l2h : l2h - l2w;
    if (l2w > l2h && mode < 2 + shift[d]) return mode + 65;
    if (l2h > l2w && mode > 66 - shift[d]) return mode - 65 - (account_for_dc_planar ? 2 : 0);
  }
  return mode;
}

/* Position of LFNST scan entry j in a block of width w: up-right diagonal order inside a 4x4
 * group and over the 2x2 groups of the top-left 8x8 (uvg_coef_top_left_diag_scan_8x8,
 * lfnst_tables.h:366; first group of uvg_g_sig_last_scan[SCAN_DIAG] for narrower blocks). */
static int lfnst_scan_pos(int j, int w, int whge3)
{
  static const uint8_t dx[16] = { 0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 1, 2, 3, 2, 3, 3 };
  static const uint8_t dy[16] = { 0, 1, 0, 2, 1, 0, 3, 2, 1, 0, 3, 2, 1, 3, 2, 3 };
  static const uint8_t gx[4] = { 0, 0, 1, 1 }, gy[4] = { 0, 1, 0, 1 };
  const int g = whge3 ? j >> 4 : 0, i = j & 15;
  return (gy[g] * 4 + dy[i]) * w + gx[g] * 4 + dx[i];
}

typedef struct { int mode, transpose, sb, zero_out, whge3; } lfnst_ctx;

/* the part uvg_fwd_lfnst and uvg_inv_lfnst share (:986-1018, :1128-1157). intra_mode: the block's
 * mode after the CCLM / MIP substitutions; wa_log2_*: cur_cu->log2_* for luma, the block's for chroma */
static lfnst_ctx lfnst_setup(int w, int h, int intra_mode, int wa_log2_w, int wa_log2_h)
{
  lfnst_ctx c;
  int m = wide_angle(intra_mode, wa_log2_w, wa_log2_h, 1);
  m = m < 0 ? m + 14 + 67 : m >= 67 ? m + 14 : m;                 /* get_lfnst_intra_mode :932-950 */
  c.mode = m;
  c.transpose = (m >= 67 && m >= 81) || (m < 67 && m > 34);       /* get_transpose_flag :952-956 */
  c.whge3 = w >= 8 && h >= 8;
  c.sb = c.whge3 ? 8 : 4;
  c.zero_out = ((w == 4 && h == 4) || (w == 8 && h == 8)) ? 8 : 16;
  return c;
}

void orc_fwd_lfnst(int w, int h, int lfnst_idx, int intra_mode, int wa_log2_w, int wa_log2_h, int16_t *coeffs)
{
  const lfnst_ctx c = lfnst_setup(w, h, intra_mode, wa_log2_w, wa_log2_h);
  int16_t in[48], out[48];
  /* gather the top-left region, row-major or transposed (:1030-1066): 8x8 blocks use the L-shaped 48 */
  int n = 0;
  if (!c.transpose) {
    for (int y = 0; y < c.sb; ++y) {
      const int len = y < 4 ? c.sb : 4;
      for (int x = 0; x < len; ++x) in[n++] = coeffs[y * w + x];
    }
  } else if (c.sb == 4) {
    for (int y = 0; y < 4; ++y) for (int x = 0; x < 4; ++x) in[4 * x + y] = coeffs[y * w + x];
  } else {
    for (int y = 0; y < 8; ++y) {
      for (int x = 0; x < 4; ++x) in[8 * x + y] = coeffs[y * w + x];
      if (y < 4) for (int x = 4; x < 8; ++x) in[32 + 4 * (x - 4) + y] = coeffs[y * w + x];
    }
  }
  /* uvg_fwd_lfnst_NxN :878-915 */
  const int tr_size = c.sb > 4 ? 48 : 16;
  const int8_t *mat = c.sb > 4 ? g_lfnst8 + ((size_t)(lfnst_set(c.mode) * 2 + lfnst_idx - 1) * 16) * 48
                               : g_lfnst4 + ((size_t)(lfnst_set(c.mode) * 2 + lfnst_idx - 1) * 16) * 16;
  for (int j = 0; j < tr_size; ++j) {
    int acc = 0;
    if (j < c.zero_out) {
      for (int i = 0; i < tr_size; ++i) acc += in[i] * mat[j * tr_size + i];
      out[j] = (int16_t)((acc + 64) >> 7);
    } else {
      out[j] = 0;
    }
  }
  for (int j = 0; j < tr_size; ++j) coeffs[lfnst_scan_pos(j, w, c.whge3)] = out[j];    /* :1071-1081 */
}

void orc_inv_lfnst(int w, int h, int lfnst_idx, int intra_mode, int wa_log2_w, int wa_log2_h, int16_t *coeffs)
{
  const lfnst_ctx c = lfnst_setup(w, h, intra_mode, wa_log2_w, wa_log2_h);
  int16_t in[16], out[48];
  for (int j = 0; j < 16; ++j) in[j] = coeffs[lfnst_scan_pos(j, w, c.whge3)];          /* :1165-1169 */
  /* uvg_inv_lfnst_NxN :1086-1108: the cast to coeff_t comes before the clip, which is then a no-op */
  const int tr_size = c.sb > 4 ? 48 : 16;
  const int8_t *mat = c.sb > 4 ? g_lfnst8 + ((size_t)(lfnst_set(c.mode) * 2 + lfnst_idx - 1) * 16) * 48
                               : g_lfnst4 + ((size_t)(lfnst_set(c.mode) * 2 + lfnst_idx - 1) * 16) * 16;
  for (int j = 0; j < tr_size; ++j) {
    int acc = 0;
    for (int i = 0; i < c.zero_out; ++i) acc += in[i] * mat[i * tr_size + j];
    out[j] = (int16_t)((acc + 64) >> 7);
  }
  int n = 0;
  if (!c.transpose) {                                                                   /* :1175-1198 */
    for (int y = 0; y < c.sb; ++y) {
      const int len = y < 4 ? c.sb : 4;
      for (int x = 0; x < len; ++x) coeffs[y * w + x] = out[n++];
    }
  } else if (c.sb == 4) {
    for (int y = 0; y < 4; ++y) for (int x = 0; x < 4; ++x) coeffs[y * w + x] = out[4 * x + y];
  } else {
    for (int y = 0; y < 8; ++y) {
      for (int x = 0; x < 4; ++x) coeffs[y * w + x] = out[8 * x + y];
      if (y < 4) for (int x = 4; x < 8; ++x) coeffs[y * w + x] = out[32 + 4 * (x - 4) + y];
    }
  }
}

/* uvg_quantize_residual_generic with cfg.lfnst (quant-generic.c:497-566), plain quant path.
 * li: what the composite reads from cu_info_t / the state for LFNST. */
int orc_quantize_residual_lfnst(const orc_quant_params *p, const orc_lfnst_info *li, int cu_type, int tr_idx,
                                int w, int h, int color, int in_stride, int out_stride,
                                const orc_pixel *ref_in, const orc_pixel *pred_in, orc_pixel *rec_out,
                                int16_t *coeff_out, int early_skip)
{
  int16_t residual[32 * 32], coeff[32 * 32];
  orc_generate_residual(ref_in, pred_in, residual, w, h, in_stride, in_stride);
  if (p->mts || w != h) orc_mts_dct(p->bitdepth, color, cu_type, tr_idx, li->lfnst_idx, li->cr_lfnst_idx, 0, w, h, residual, coeff, p->mts);
  else orc_dct_nxn(w, p->bitdepth, residual, coeff);
  const int lfnst_index = (li->tree_type != 2 || color == 0) ? li->lfnst_idx : li->cr_lfnst_idx;   /* :509 */
  const int ts_flag = (tr_idx == 1 && color == 0);
  /* uvg_fwd_lfnst's own gate (:983-990) */
  const int sep_tree = li->cu_log2_w + li->cu_log2_h < 6 || li->tree_type != 0;
  int mode = color == 0 ? li->intra_mode : li->mode_chroma;
  if (mode >= 81 && mode <= 83) mode = li->luma_mode;
  const int l2w = ilog2(w), l2h = ilog2(h);
  const int wl2w = color == 0 ? li->cu_log2_w : l2w, wl2h = color == 0 ? li->cu_log2_h : l2h;
  const int do_lfnst = cu_type == 1 && lfnst_index && !ts_flag && (color == 0 || sep_tree);
  if (do_lfnst) orc_fwd_lfnst(w, h, lfnst_index, mode, wl2w, wl2h, coeff);
  orc_quant(p, coeff, coeff_out, w, h, color, cu_type, ts_flag, lfnst_index);
  int has_coeffs = 0;
  for (int i = 0; i < w * h; ++i) if (coeff_out[i] != 0) { has_coeffs = 1; break; }
  if (has_coeffs && !early_skip) {
    orc_dequant(p, coeff_out, coeff, w, h, color, cu_type, ts_flag);
    if (do_lfnst) orc_inv_lfnst(w, h, lfnst_index, mode, wl2w, wl2h, coeff);
    if (p->mts || w != h) orc_mts_idct(p->bitdepth, color, cu_type, tr_idx, li->lfnst_idx, li->cr_lfnst_idx, 0, w, h, coeff, residual, p->mts);
    else orc_idct_nxn(w, p->bitdepth, coeff, residual);
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        int16_t val = (int16_t)(residual[x + y * w] + pred_in[x + y * in_stride]);
        rec_out[x + y * out_stride] = (orc_pixel)clip3(0, PIXEL_MAX, val);
      }
  } else if (rec_out != pred_in) {
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) rec_out[x + y * out_stride] = pred_in[x + y * in_stride];
  }
  return has_coeffs;
}

/* coeff_abs_sum_generic :671-678 */
uint32_t orc_coeff_abs_sum(const int16_t *coeffs, size_t length)
{
  uint32_t sum = 0;
  for (size_t i = 0; i < length; ++i) sum += (uint32_t)iabs(coeffs[i]);
  return sum;
}

/* fast_coeff_cost_generic :688-705: 4 Q8.8 weights packed in a uint64. */
uint32_t orc_fast_coeff_cost(const int16_t *coeff, int w, int h, uint64_t weights)
{
  uint32_t sum = 0;
  for (int i = 0; i < w * h; ++i) {
    uint32_t a = (uint32_t)iabs(coeff[i]);
    if (a > 3) a = 3;
    sum += (uint32_t)((weights >> (16 * a)) & 0xffff);
  }
  return (sum + (1 << 7)) >> 8;
}

/* ===================================================================================
 * strategies-intra: luma intra prediction of a w*h block from its
 * reference rows, as the rough search calls it (search_intra.c:1043-1046 ->
 * uvg_intra_predict intra.c:1372 -> intra_predict_regular :660), for multi_ref_idx = 0,
 * no ISP, no MIP. ref_top / ref_left follow uvg_intra_ref: index 0 is the top-left corner,
 * index 1.. the row above / column left of the block, 2*w+1 (2*h+1) samples.
 * ================================================================================= */
static const int16_t k_modedisp2sampledisp[32] = { 0, 1, 2, 3, 4, 6, 8, 10, 12, 14, 16, 18, 20, 23, 26, 29, 32,
  35, 39, 45, 51, 57, 64, 73, 86, 102, 128, 171, 256, 341, 512, 1024 };
static const int16_t k_modedisp2invsampledisp[32] = { 0, 16384, 8192, 5461, 4096, 2731, 2048, 1638, 1365, 1170,
  1024, 910, 819, 712, 630, 565, 512, 468, 420, 364, 321, 287, 256, 224, 191, 161, 128, 96, 64, 48, 32, 16 };
static const int8_t k_pre_scale[32] = { 8, 7, 6, 5, 5, 4, 4, 4, 3, 3, 3, 3, 3, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1,
  0, 0, 0, -1, -1, -2, -3 };
static const int8_t k_hor_ver_dist_thres[8] = { 24, 24, 24, 14, 2, 0, 0, 0 };
/* VVC 4-tap cubic intra interpolation filter (intra-generic.c:77-111), rows = 1/32 phase */
static const int8_t k_cubic[32][4] = {
  { 0, 64, 0, 0 }, { -1, 63, 2, 0 }, { -2, 62, 4, 0 }, { -2, 60, 7, -1 }, { -2, 58, 10, -2 }, { -3, 57, 12, -2 },
  { -4, 56, 14, -2 }, { -4, 55, 15, -2 }, { -4, 54, 16, -2 }, { -5, 53, 18, -2 }, { -6, 52, 20, -2 },
  { -6, 49, 24, -3 }, { -6, 46, 28, -4 }, { -5, 44, 29, -4 }, { -4, 42, 30, -4 }, { -4, 39, 33, -4 },
  { -4, 36, 36, -4 }, { -4, 33, 39, -4 }, { -4, 30, 42, -4 }, { -4, 29, 44, -5 }, { -4, 28, 46, -6 },
  { -3, 24, 49, -6 }, { -2, 20, 52, -6 }, { -2, 18, 53, -5 }, { -2, 16, 54, -4 }, { -2, 15, 55, -4 },
  { -2, 14, 56, -4 }, { -2, 12, 57, -3 }, { -2, 10, 58, -2 }, { -1, 7, 60, -2 }, { 0, 4, 62, -2 }, { 0, 2, 63, -1 } };

/* uvg_angular_pred_generic, intra-generic.c:55-299, luma, multi_ref_idx = 0, isp = 0. */
/* The strategy leaf uvg_angular_pred_generic (intra-generic.c:55-299) with all its arguments:
 * w x h = the block in the plane that is predicted (chroma dims for channel_type != 0),
 * intra_mode already wide-angle corrected (may be < 2 or > 66), multi_ref_idx, isp_mode. */
void orc_angular_pred_leaf(int w, int h, int intra_mode, int channel_type, const orc_pixel *ref_above,
                           const orc_pixel *ref_left, orc_pixel *dst, int mrl, int isp)
{
  const int l2w = ilog2(w), l2h = ilog2(h);
  const int pred_mode = intra_mode;
  const int vertical = intra_mode >= 34;
  const int mode_disp = vertical ? pred_mode - 50 : 18 - pred_mode;
  const int ad = mode_disp < 0 ? -mode_disp : mode_disp;
  const int sample_disp = (mode_disp < 0 ? -1 : 1) * k_modedisp2sampledisp[ad];
  const int side_size = vertical ? l2h : l2w;
  int scale = side_size - k_pre_scale[ad];
  if (scale > 2) scale = 2;
  orc_pixel main_buf[2 * 128 + 3 + 33 * 3 + 128], *ref_main;
  const orc_pixel *ref_side;
  orc_pixel work[64 * 64];
  int W = w, H = h;
  if (sample_disp < 0) {
    /* project the side reference onto the negative part of the main reference (:145-156) */
    const orc_pixel *m = vertical ? ref_above : ref_left, *sd = vertical ? ref_left : ref_above;
    const int msz = vertical ? w : h, ssz = vertical ? h : w;
    ref_main = main_buf + 128;
    for (int i = 0; i < msz + 2 + mrl; ++i) ref_main[i] = m[i];
    for (int i = -ssz; i <= -1; ++i) {
      int idx = (-i * k_modedisp2invsampledisp[ad] + 256) >> 9;
      if (idx > ssz) idx = ssz;
      ref_main[i] = sd[idx];
    }
    ref_side = sd;
  } else {
    ref_main = (orc_pixel *)(vertical ? ref_above : ref_left);
    ref_side = vertical ? ref_left : ref_above;
  }
  ref_main += mrl;                                     /* line offset in the reference buffers (:158-160) */
  ref_side += mrl;
  if (!vertical) { int t = W; W = H; H = t; }
  if (sample_disp != 0) {
    int use_cubic = 1;
    const int thr = k_hor_ver_dist_thres[(l2w + l2h) >> 1];
    int d50 = pred_mode - 50, d18 = pred_mode - 18;
    if (d50 < 0) d50 = -d50;
    if (d18 < 0) d18 = -d18;
    const int dist = d50 < d18 ? d50 : d18;
    const int frac_mode = ((sample_disp < 0 ? -sample_disp : sample_disp) & 0x1f) != 0;
    if (dist > thr && frac_mode) use_cubic = 0;
    if (mrl || isp) use_cubic = 1;                     /* :176-179 */
    for (int y = 0, delta_pos = sample_disp * (1 + mrl); y < H; ++y, delta_pos += sample_disp) {
      const int delta_int = delta_pos >> 5, fr = delta_pos & 31;
      const int gauss[4] = { 16 - (fr >> 1), 32 - (fr >> 1), 16 + (fr >> 1), fr >> 1 };
      if (frac_mode) {
        if (channel_type == 0) {
          for (int x = 0; x < W; ++x) {
            const orc_pixel *p = &ref_main[delta_int + x];
            int s = 32;
            for (int k = 0; k < 4; ++k) s += (use_cubic ? k_cubic[fr][k] : gauss[k]) * (int)p[k];
            work[y * W + x] = (orc_pixel)clip3(0, PIXEL_MAX, s >> 6);
          }
        } else {                                       /* chroma: 2-tap linear (:207-214) */
          for (int x = 0; x < W; ++x) {
            const int r1 = ref_main[x + delta_int + 1], r2 = ref_main[x + delta_int + 2];
            work[y * W + x] = (orc_pixel)(r1 + ((fr * (r2 - r1) + 16) >> 5));
          }
        }
      } else {
        for (int x = 0; x < W; ++x) work[y * W + x] = ref_main[x + delta_int + 1];
      }
      /* PDPC for angular modes (:226-251); wide-angle modes (outside 2..66) skip the mode tests */
      int pdpc = (W >= 4 && H >= 4) && mrl == 0;
      if (pred_mode > 1 && pred_mode < 67) {
        if (mode_disp < 0 || mrl) pdpc = 0;
        else if (mode_disp > 0) pdpc = pdpc && (scale >= 0);
      }
      if (pdpc && scale >= 0) {
        int inv_angle_sum = 256;
        const int lim = (3 << scale) < W ? (3 << scale) : W;
        for (int x = 0; x < lim; ++x) {
          inv_angle_sum += k_modedisp2invsampledisp[ad];
          const int wL = 32 >> ((2 * x) >> scale);
          const int left = ref_side[y + (inv_angle_sum >> 9) + 1];
          const int cur = work[y * W + x];
          work[y * W + x] = (orc_pixel)(cur + ((wL * (left - cur) + 32) >> 6));
        }
      }
    }
  } else {
    /* pure horizontal / vertical (:254-280) with PDPC */
    const int do_pdpc = (W >= 4 && H >= 4) && mrl == 0;
    const int sc = (l2w + l2h - 2) >> 2;
    const int top_left = ref_main[0];
    for (int y = 0; y < H; ++y) {
      for (int x = 0; x < W; ++x) work[y * W + x] = ref_main[1 + x];
      if (do_pdpc) {
        const int left = ref_side[1 + y];
        const int lim = (3 << sc) < W ? (3 << sc) : W;
        for (int x = 0; x < lim; ++x) {
          const int wL = 32 >> ((2 * x) >> sc);
          work[y * W + x] = (orc_pixel)clip3(0, PIXEL_MAX, (int)work[y * W + x] + ((wL * (left - top_left) + 32) >> 6));
        }
      }
    }
  }
  if (vertical) {
    for (int i = 0; i < w * h; ++i) dst[i] = work[i];
  } else {
    /* horizontal modes were predicted transposed (:283-298) */
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) dst[y * w + x] = work[x * W + y];
  }
}

/* luma, reference line 0, no ISP */
void orc_angular_pred(int w, int h, int pred_mode, const orc_pixel *ref_above, const orc_pixel *ref_left,
                      orc_pixel *dst)
{
  orc_angular_pred_leaf(w, h, pred_mode, 0, ref_above, ref_left, dst, 0, 0);
}

/* uvg_intra_pred_planar_generic :306-360 */
void orc_intra_pred_planar(int w, int h, const orc_pixel *ref_top, const orc_pixel *ref_left, orc_pixel *dst)
{
  const int l2w = ilog2(w), l2h = ilog2(h);
  const int top_right = ref_top[w + 1], bottom_left = ref_left[h + 1];
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const int hor = ((int)ref_left[y + 1] << l2w) + (x + 1) * (top_right - (int)ref_left[y + 1]);
      const int ver = ((int)ref_top[x + 1] << l2h) + (y + 1) * (bottom_left - (int)ref_top[x + 1]);
      dst[y * w + x] = (orc_pixel)(((hor << l2h) + (ver << l2w) + (1 << (l2w + l2h))) >> (1 + l2w + l2h));
    }
}

/* intra_pred_dc, intra.c:236-273 (multi_ref_idx = 0) */
void orc_intra_pred_dc(int w, int h, const orc_pixel *ref_top, const orc_pixel *ref_left, orc_pixel *dst)
{
  int sum = 0;
  if (w >= h) for (int i = 0; i < w; ++i) sum += ref_top[i + 1];
  if (w <= h) for (int j = 0; j < h; ++j) sum += ref_left[j + 1];
  const int denom = w == h ? w << 1 : (w > h ? w : h);
  const int dc = (sum + (denom >> 1)) >> ilog2(denom);
  for (int i = 0; i < w * h; ++i) dst[i] = (orc_pixel)dc;
}

/* uvg_pdpc_planar_dc_generic :414-437 */
void orc_pdpc_planar_dc(int w, int h, const orc_pixel *ref_top, const orc_pixel *ref_left, orc_pixel *dst)
{
  const int scale = (ilog2(w) + ilog2(h) - 2) >> 2;
  for (int y = 0; y < h; ++y) {
    int sy = (y << 1) >> scale; if (sy > 31) sy = 31;
    const int wT = 32 >> sy;
    for (int x = 0; x < w; ++x) {
      int sx = (x << 1) >> scale; if (sx > 31) sx = 31;
      const int wL = 32 >> sx;
      const int cur = dst[x + y * w];
      dst[x + y * w] = (orc_pixel)(cur + ((wL * ((int)ref_left[y + 1] - cur) + wT * ((int)ref_top[x + 1] - cur) + 32) >> 6));
    }
  }
}

/* intra_filter_reference, intra.c:190-225: [1 2 1]/4 smoothing of the reference rows */
void orc_intra_filter_reference(int w, int h, const orc_pixel *top, const orc_pixel *left, orc_pixel *ftop,
                                orc_pixel *fleft)
{
  const int rw = 2 * w + 1, rh = 2 * h + 1;
  fleft[0] = (orc_pixel)((left[1] + 2 * left[0] + top[1] + 2) >> 2);
  ftop[0] = fleft[0];
  for (int y = 1; y < rh - 1; ++y) fleft[y] = (orc_pixel)((left[y - 1] + 2 * left[y] + left[y + 1] + 2) >> 2);
  fleft[rh - 1] = left[rh - 1];
  for (int x = 1; x < rw - 1; ++x) ftop[x] = (orc_pixel)((top[x - 1] + 2 * top[x] + top[x + 1] + 2) >> 2);
  ftop[rw - 1] = top[rw - 1];
}

/* intra_predict_regular, intra.c:660-753, luma, square or not, mrl = 0, isp = 0,
 * intra_smoothing_disabled = 0: reference-smoothing decision + predictor + PDPC. */
void orc_intra_predict(int w, int h, int mode, const orc_pixel *top, const orc_pixel *left, orc_pixel *dst)
{
  const int l2w = ilog2(w), l2h = ilog2(h);
  int pred_mode = mode;
  if (l2w != l2h && mode > 1 && mode <= 66) {                /* uvg_wide_angle_correction :637 */
    static const int shift[] = { 0, 6, 10, 12, 14, 15 };
    const int ds = l2w > l2h ? l2w - l2h : l2h - l2w;
    if (l2w > l2h && mode < 2 + shift[ds]) pred_mode += 65;
    else if (l2h > l2w && mode > 66 - shift[ds]) pred_mode -= 65;
  }
  int filtered = 0;
  if (mode == 1 || (w == 4 && h == 4)) {
  } else if (mode == 0) {
    filtered = w * h > 32;
  } else {
    const int thr = k_hor_ver_dist_thres[(l2w + l2h) >> 1];
    int d50 = pred_mode - 50, d18 = pred_mode - 18;
    if (d50 < 0) d50 = -d50;
    if (d18 < 0) d18 = -d18;
    if ((d50 < d18 ? d50 : d18) > thr) {
      const int md = pred_mode >= 34 ? pred_mode - 50 : 18 - pred_mode;
      const int sd = k_modedisp2sampledisp[md < 0 ? -md : md];
      if ((sd & 0x1f) == 0) filtered = 1;
    }
  }
  /* The encoder's reference rows continue past their 2n+1 samples with the last sample
   * replicated (uvg_intra_build_reference_any, intra.c:879-886, :1044-1051): the 4-tap filter
   * of the steepest angles reads up to 2 samples beyond 2n. */
  enum { PAD = 8 };
  orc_pixel tp[2 * 64 + 1 + PAD], lp[2 * 64 + 1 + PAD], ftop[2 * 64 + 1 + PAD], fleft[2 * 64 + 1 + PAD];
  for (int i = 0; i < 2 * w + 1 + PAD; ++i) tp[i] = top[i < 2 * w + 1 ? i : 2 * w];
  for (int i = 0; i < 2 * h + 1 + PAD; ++i) lp[i] = left[i < 2 * h + 1 ? i : 2 * h];
  const orc_pixel *ut = tp, *ul = lp;
  if (filtered) {
    orc_intra_filter_reference(w, h, tp, lp, ftop, fleft);
    for (int i = 2 * w + 1; i < 2 * w + 1 + PAD; ++i) ftop[i] = ftop[2 * w];
    for (int i = 2 * h + 1; i < 2 * h + 1 + PAD; ++i) fleft[i] = fleft[2 * h];
    ut = ftop; ul = fleft;
  }
  if (mode == 0) orc_intra_pred_planar(w, h, ut, ul, dst);
  else if (mode == 1) orc_intra_pred_dc(w, h, ut, ul, dst);
  else orc_angular_pred(w, h, pred_mode, ut, ul, dst);
  if ((mode == 0 || mode == 1) && w >= 4 && h >= 4) orc_pdpc_planar_dc(w, h, ut, ul, dst);
}

/* =====================================================================================
 * strategies-ipol (next row 8(f)-1): motion-compensation sample interpolation,
 * generic/ipol-generic.c:134-210 (luma, 8 taps, 1/16 phases) and :681-759 (chroma, 4 taps,
 * 1/32 phases). `src` points at the integer sample of the block's top-left; the filters reach 3
 * (1) samples up/left and 4 (2) down/right. Horizontal pass into int16 (>> BD-8), vertical pass
 * (>> 6), then either the uni-prediction rounding to a pixel (>> 14-BD with offset, clipped) or,
 * for the "_hi" variants, the 14-bit intermediate itself.
 * ===================================================================================== */
/* VVC luma interpolation filter, phases 0..8 of 16 (filter.c:62-80); phases 9..15 mirror 7..1 */
static const int8_t k_luma_half[9][8] = {
  { 0, 0, 0, 64, 0, 0, 0, 0 },     { 0, 1, -3, 63, 4, -2, 1, 0 },    { -1, 2, -5, 62, 8, -3, 1, 0 },
  { -1, 3, -8, 60, 13, -4, 1, 0 }, { -1, 4, -10, 58, 17, -5, 1, 0 }, { -1, 4, -11, 52, 26, -8, 3, -1 },
  { -1, 3, -9, 47, 31, -10, 4, -1 }, { -1, 4, -11, 45, 34, -10, 4, -1 }, { -1, 4, -11, 40, 40, -11, 4, -1 } };
static void luma_filter(int phase, int f[8])
{
  for (int i = 0; i < 8; ++i) f[i] = phase <= 8 ? k_luma_half[phase][i] : k_luma_half[16 - phase][7 - i];
}

static void orc_sample(int taps, int w, int h, const orc_pixel *src, int src_stride, const int *fh, const int *fv,
                       orc_pixel *dst, int16_t *dst_hi, int dst_stride)
{
  const int off = taps == 8 ? 3 : 1, pad = taps - 1;
  const int shift1 = BD - 8, shift2 = 6, wp_shift = 14 - BD, wp_off = 1 << (wp_shift - 1);
  static __thread int16_t hor[(64 + 7) * 64];
  for (int y = 0; y < h + pad; ++y)
    for (int x = 0; x < w; ++x) {
      int s = 0;
      for (int i = 0; i < taps; ++i) s += fh[i] * (int)src[(y - off) * src_stride + (x - off) + i];
      hor[y * 64 + x] = (int16_t)(s >> shift1);
    }
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int s = 0;
      for (int i = 0; i < taps; ++i) s += fv[i] * (int)hor[(y + i) * 64 + x];
      s >>= shift2;
      if (dst_hi) dst_hi[y * dst_stride + x] = (int16_t)s;
      else dst[y * dst_stride + x] = (orc_pixel)clip3(0, PIXEL_MAX, (s + wp_off) >> wp_shift);
    }
}

void orc_sample_quarterpel_luma(int w, int h, const orc_pixel *src, int src_stride, orc_pixel *dst, int dst_stride,
                                int mvx, int mvy)
{
  int fh[8], fv[8];
  luma_filter(mvx & 15, fh); luma_filter(mvy & 15, fv);
  orc_sample(8, w, h, src, src_stride, fh, fv, dst, NULL, dst_stride);
}
void orc_sample_quarterpel_luma_hi(int w, int h, const orc_pixel *src, int src_stride, int16_t *dst, int dst_stride,
                                   int mvx, int mvy)
{
  int fh[8], fv[8];
  luma_filter(mvx & 15, fh); luma_filter(mvy & 15, fv);
  orc_sample(8, w, h, src, src_stride, fh, fv, NULL, dst, dst_stride);
}
void orc_sample_octpel_chroma(int w, int h, const orc_pixel *src, int src_stride, orc_pixel *dst, int dst_stride,
                              int mvx, int mvy)
{
  int fh[4], fv[4];
  for (int i = 0; i < 4; ++i) { fh[i] = k_cubic[mvx & 31][i]; fv[i] = k_cubic[mvy & 31][i]; }   /* filter.c:82-116 = the intra cubic table */
  orc_sample(4, w, h, src, src_stride, fh, fv, dst, NULL, dst_stride);
}
void orc_sample_octpel_chroma_hi(int w, int h, const orc_pixel *src, int src_stride, int16_t *dst, int dst_stride,
                                 int mvx, int mvy)
{
  int fh[4], fv[4];
  for (int i = 0; i < 4; ++i) { fh[i] = k_cubic[mvx & 31][i]; fv[i] = k_cubic[mvy & 31][i]; }
  orc_sample(4, w, h, src, src_stride, fh, fv, NULL, dst, dst_stride);
}

/* =====================================================================================
 * The two remaining non-distortion entries of the picture strategy table (SURVEY 8(a) "next"):
 * crc32c_4x4 / crc32c_8x8 (picture-generic.c:1371-1443: CRC-32C, reflected polynomial 0x82F63B78,
 * over the block's pixel BYTES row by row - little-endian byte pairs for > 8-bit pixels) and
 * bipred_average (:1133-1247: average of two predictions, each either pixels or 14-bit
 * intermediates). pixel_var (:1334-1357) accumulates in double in index order and is therefore
 * not reproducible by a parallel reduction; it stays with the reference.
 * ===================================================================================== */
uint32_t orc_crc32c_nxn(int n, const orc_pixel *buf, uint32_t stride)
{
  static uint32_t table[256];
  if (!table[1])
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int b = 0; b < 8; ++b) c = (c >> 1) ^ ((c & 1) ? 0x82F63B78u : 0);
      table[i] = c;
    }
  uint32_t crc = 0xFFFFFFFFu;
  for (int y = 0; y < n; ++y) {
    const uint8_t *row = (const uint8_t *)(buf + (size_t)y * stride);
    for (size_t i = 0; i < (size_t)n * sizeof(orc_pixel); ++i) crc = (crc >> 8) ^ table[(crc ^ row[i]) & 0xFF];
  }
  return crc ^ 0xFFFFFFFFu;
}

void orc_bipred_average(orc_pixel *dst, int dst_stride, const void *src0, int src0_is_im, const void *src1,
                        int src1_is_im, int w, int h)
{
  const int shift = 15 - BD, offset = 1 << (shift - 1);
  for (int i = 0; i < w * h; ++i) {
    const int16_t a = src0_is_im ? ((const int16_t *)src0)[i] : (int16_t)(((const orc_pixel *)src0)[i] << (14 - BD));
    const int16_t b = src1_is_im ? ((const int16_t *)src1)[i] : (int16_t)(((const orc_pixel *)src1)[i] << (14 - BD));
    dst[(i / w) * dst_stride + (i % w)] = (orc_pixel)clip3(0, PIXEL_MAX, ((int)a + (int)b + offset) >> shift);
  }
}

/* The four FME block filters (ipol-generic.c:213-680, uvg_filter_hpel_blocks_hor_ver / _diag,
 * uvg_filter_qpel_blocks_hor_ver / _diag) share buffers of horizontally filtered rows between calls,
 * but each of their four output blocks is exactly the sample interpolation of the block at one
 * quarter-pel offset from its integer position (which is sample (1,1) of `src`):
 *   step 0: (-2,0) (2,0) (0,-2) (0,2)         step 1: (-2,-2) (2,-2) (-2,2) (2,2)
 *   step 2: c+(-1,0) c+(1,0) c+(0,-1) c+(0,1)  step 3: c+(-1,-1) c+(1,-1) c+(-1,1) c+(1,1),  c = 2*(off_x, off_y)
 * (established against the reference by tests/test_oracle_ipol_vs_ref.py). out: dense [4][h][w]. */
void orc_fme_filter(int step, const orc_pixel *src, int src_stride, int w, int h, int off_x, int off_y, orc_pixel *out)
{
  static const int8_t d[4][4][2] = { { { -2, 0 }, { 2, 0 }, { 0, -2 }, { 0, 2 } }, { { -2, -2 }, { 2, -2 }, { -2, 2 }, { 2, 2 } },
                                     { { -1, 0 }, { 1, 0 }, { 0, -1 }, { 0, 1 } }, { { -1, -1 }, { 1, -1 }, { -1, 1 }, { 1, 1 } } };
  for (int j = 0; j < 4; ++j) {
    const int qx = (step >= 2 ? 2 * off_x : 0) + d[step][j][0], qy = (step >= 2 ? 2 * off_y : 0) + d[step][j][1];
    const int mvx = qx * 4, mvy = qy * 4;               /* 1/16 samples */
    const orc_pixel *p = src + (1 + (mvy >> 4)) * src_stride + 1 + (mvx >> 4);
    orc_sample_quarterpel_luma(w, h, p, src_stride, out + (size_t)j * w * h, w, mvx, mvy);
  }
}

/* pixel_var_generic, picture-generic.c:1334-1357: mean, then squared deviations accumulated in index order. */
double orc_pixel_var(const orc_pixel *arr, uint32_t len)
{
  double sum = 0, var = 0;
  for (uint32_t i = 0; i < len; ++i) sum += arr[i];
  const double mean = sum / (double)len;
  for (uint32_t i = 0; i < len; ++i) {
    volatile double tmp = (double)arr[i] - mean;      /* two roundings: the difference, then the product */
    var += tmp * tmp;
  }
  var /= len;
  return var;
}
