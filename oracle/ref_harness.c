/*
 * ref_harness.c - TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin plain-C access layer over the UNMODIFIED reference (compiled from
 * /root/reference/src by oracle/Makefile into oracle/_ref/bd{8,10}/libuvg266ref.so).
 * It exists because some reference leaf functions take reference-private structs
 * (encoder_state_t, cu_info_t) that ctypes cannot build; the wrappers below build them
 * the way the reference's own tests do:
 *   - strategy collection:  tests/test_strategies.c:39-65
 *   - cu_info_t for mts:     tests/mts_tests.c:109-114
 *   - encoder state:         uvg266.c:90 (uvg266_open) + uvg266_internal.h:50-53
 * It is compiled against the reference headers where they lie; no reference source
 * is copied.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sched.h>
#include <pthread.h>
#include <time.h>
#include <unistd.h>

#include "global.h"
#include "uvg266.h"
#include "uvg266_internal.h"
#include "encoderstate.h"
#include "encoder.h"
#include "cu.h"
#include "strategyselector.h"
#include "strategies/strategies-picture.h"
#include "strategies/strategies-dct.h"
#include "strategies/strategies-quant.h"
#include "tables.h"
#include "intra.h"
#include "strategies/strategies-intra.h"
#include "strategies/strategies-ipol.h"

static strategy_list_t g_list;
static const uvg_api *g_api;
static uvg_config *g_cfg;
static uvg_encoder *g_enc;
static encoder_state_t *g_state;

int refh_bitdepth(void) { return UVG_BIT_DEPTH; }
int refh_pixel_size(void) { return (int)sizeof(uvg_pixel); }

/* Look up one registered implementation, e.g. ("satd_8x8", "generic"). */
void *refh_get(const char *type, const char *strategy)
{
  for (unsigned i = 0; i < g_list.count; ++i) {
    if (!strcmp(g_list.strategies[i].type, type) &&
        !strcmp(g_list.strategies[i].strategy_name, strategy)) {
      return g_list.strategies[i].fptr;
    }
  }
  return NULL;
}

/* Point every hot-path global (uvg_reg_sad, uvg_quant, ...) at the named strategy when it
 * has one for that type (else keep the current). Composite functions (quantize_residual)
 * call through these globals, so "generic" makes the whole chain the C oracle. */
int refh_use(const char *strategy)
{
  int n = 0;
  for (const strategy_to_select_t *s = strategies_to_select; s->fptr; ++s) {
    void *f = refh_get(s->strategy_type, strategy);
    if (f) { *(s->fptr) = f; ++n; }
  }
  return n;
}

/* Open an encoder (64x64 dummy frame) with "key=value,key=value" options, exactly through
 * the public API, then collect all picture/dct/quant strategies and select `generic`. */
int refh_open(const char *opts)
{
  if (g_enc) { g_api->encoder_close(g_enc); g_enc = NULL; }
  if (g_cfg) { g_api->config_destroy(g_cfg); g_cfg = NULL; }
  g_api = uvg_api_get(UVG_BIT_DEPTH);
  if (!g_api) return 0;
  g_cfg = g_api->config_alloc();
  if (!g_cfg || !g_api->config_init(g_cfg)) return 0;
  g_cfg->width = 64; g_cfg->height = 64;
  g_cfg->threads = 0; g_cfg->owf = 0; g_cfg->wpp = 0;
  g_cfg->cpuid = 1;
  char *dup = strdup(opts ? opts : "");
  char *save = NULL; /* strtok_r: config_parse itself uses strtok for presets */
  for (char *tok = strtok_r(dup, ",", &save); tok; tok = strtok_r(NULL, ",", &save)) {
    char *eq = strchr(tok, '=');
    const char *val = "true";
    if (eq) { *eq = 0; val = eq + 1; }
    if (!g_api->config_parse(g_cfg, tok, val)) {
      fprintf(stderr, "refh_open: bad option %s=%s\n", tok, val);
      free(dup);
      return 0;
    }
  }
  free(dup);
  g_enc = g_api->encoder_open(g_cfg);
  if (!g_enc) return 0;
  g_state = &g_enc->states[0];

  free(g_list.strategies);
  memset(&g_list, 0, sizeof(g_list));
  if (!uvg_strategy_register_picture(&g_list, UVG_BIT_DEPTH)) return 0;
  if (!uvg_strategy_register_dct(&g_list, UVG_BIT_DEPTH)) return 0;
  if (!uvg_strategy_register_quant(&g_list, UVG_BIT_DEPTH)) return 0;
  if (!uvg_strategy_register_intra(&g_list, UVG_BIT_DEPTH)) return 0;
  if (!uvg_strategy_register_ipol(&g_list, UVG_BIT_DEPTH)) return 0;
  refh_use("generic");
  return 1;
}

void refh_set_state(int qp, int slice_is_intra)
{
  g_state->qp = (int8_t)qp;
  g_state->frame->slicetype = slice_is_intra ? UVG_SLICE_I : UVG_SLICE_P;
}

int refh_slice_is_intra(void) { return g_state->frame->slicetype == UVG_SLICE_I; }

static void fill_cu(cu_info_t *cu, int cu_type, int tr_idx, int lfnst_idx, int cr_lfnst_idx, int isp_mode)
{
  memset(cu, 0, sizeof(*cu));
  cu->type = cu_type;          /* CU_INTRA=1, CU_INTER=2 (cu.h) */
  cu->tr_idx = tr_idx;
  cu->lfnst_idx = lfnst_idx;
  cu->cr_lfnst_idx = cr_lfnst_idx;
  cu->intra.isp_mode = isp_mode;
}

/* mts_dct / mts_idct of the named strategy. */
int refh_mts(const char *strategy, int inverse, int bitdepth, int color, int cu_type,
             int tr_idx, int lfnst_idx, int cr_lfnst_idx, int isp_mode, int w, int h,
             const int16_t *in, int16_t *out, int mts_type)
{
  mts_dct_func *f = refh_get(inverse ? "mts_idct" : "mts_dct", strategy);
  if (!f) return 0;
  cu_info_t cu;
  fill_cu(&cu, cu_type, tr_idx, lfnst_idx, cr_lfnst_idx, isp_mode);
  f((int8_t)bitdepth, (color_t)color, &cu, (int8_t)w, (int8_t)h, in, out, (int8_t)mts_type);
  return 1;
}

void refh_get_tr_type(int w, int h, int color, int cu_type, int tr_idx, int lfnst_idx,
                      int cr_lfnst_idx, int isp_mode, int mts_type, int *hor, int *ver)
{
  cu_info_t cu;
  fill_cu(&cu, cu_type, tr_idx, lfnst_idx, cr_lfnst_idx, isp_mode);
  tr_type_t th, tv;
  uvg_get_tr_type((int8_t)w, (int8_t)h, (color_t)color, &cu, &th, &tv, (int8_t)mts_type);
  *hor = th; *ver = tv;
}

int refh_quant(const char *strategy, const int16_t *coef, int16_t *q_coef, int w, int h,
               int color, int scan_idx, int block_type, int transform_skip, int lfnst_idx)
{
  quant_func *f = refh_get("quant", strategy);
  if (!f) return 0;
  f(g_state, (coeff_t *)coef, q_coef, w, h, (color_t)color, (int8_t)scan_idx,
    (int8_t)block_type, (int8_t)transform_skip, (uint8_t)lfnst_idx);
  return 1;
}

int refh_dequant(const char *strategy, const int16_t *q_coef, int16_t *coef, int w, int h,
                 int color, int block_type, int transform_skip)
{
  dequant_func *f = refh_get("dequant", strategy);
  if (!f) return 0;
  f(g_state, (coeff_t *)q_coef, coef, w, h, (color_t)color, (int8_t)block_type,
    (int8_t)transform_skip);
  return 1;
}

/* cur_cu->intra.isp_mode for the next refh_quantize_residual calls (ISP + explicit MTS: DST-7) */
static int g_isp_mode;
void refh_set_isp_mode(int m) { g_isp_mode = m; }

/* quantize_residual of the named strategy; returns has_coeffs, or -1. */
int refh_quantize_residual(const char *strategy, int cu_type, int tr_idx, int lfnst_idx,
                           int w, int h, int color, int use_trskip, int in_stride,
                           int out_stride, const uvg_pixel *ref_in, const uvg_pixel *pred_in,
                           uvg_pixel *rec_out, int16_t *coeff_out, int early_skip)
{
  quant_residual_func *f = refh_get("quantize_residual", strategy);
  if (!f) return -1;
  cu_info_t cu;
  fill_cu(&cu, cu_type, tr_idx, lfnst_idx, 0, g_isp_mode);
  return (int)f(g_state, &cu, w, h, (color_t)color, SCAN_DIAG, use_trskip, in_stride,
                out_stride, ref_in, pred_in, rec_out, coeff_out, early_skip != 0, 0,
                UVG_BOTH_T);
}

/* quant_cbcr_residual (JCCR) of the named strategy; returns joint_cb_cr / 0, or -1. */
int refh_quant_cbcr_residual(const char *strategy, int cu_type, int tr_idx, int joint_cb_cr, int jccr_sign,
                             int w, int h, int in_stride, int out_stride, const uvg_pixel *u_ref,
                             const uvg_pixel *v_ref, const uvg_pixel *u_pred, const uvg_pixel *v_pred,
                             uvg_pixel *u_rec, uvg_pixel *v_rec, int16_t *coeff_out, int early_skip)
{
  quant_cbcr_func *f = refh_get("quant_cbcr_residual", strategy);
  if (!f) return -1;
  cu_info_t cu;
  fill_cu(&cu, cu_type, tr_idx, 0, 0, 0);
  cu.joint_cb_cr = (uint8_t)joint_cb_cr;
  g_state->frame->jccr_sign = jccr_sign != 0;
  return (int)f(g_state, &cu, w, h, SCAN_DIAG, in_stride, out_stride, u_ref, v_ref, u_pred, v_pred, u_rec, v_rec,
                coeff_out, early_skip != 0, 0, UVG_BOTH_T);
}

/* LFNST: the reference's kernels (exported data, lfnst_tables.h:56/:211) for the oracle's setter; the two
 * entry points themselves; quantize_residual with the LFNST fields of cu_info_t filled in. */
extern const int8_t uvg_lfnst_8x8[4][2][16][48];
extern const int8_t uvg_lfnst_4x4[4][2][16][16];
const int8_t *refh_lfnst_table(int eight) { return eight ? &uvg_lfnst_8x8[0][0][0][0] : &uvg_lfnst_4x4[0][0][0][0]; }

static void fill_cu_lfnst(cu_info_t *cu, int cu_type, int tr_idx, int lfnst_idx, int cr_lfnst_idx, int intra_mode,
                          int mode_chroma, int cu_log2_w, int cu_log2_h)
{
  fill_cu(cu, cu_type, tr_idx, lfnst_idx, cr_lfnst_idx, 0);
  cu->intra.mode = (int8_t)intra_mode;
  cu->intra.mode_chroma = (int8_t)mode_chroma;
  cu->log2_width = (uint8_t)cu_log2_w;
  cu->log2_height = (uint8_t)cu_log2_h;
}

void refh_lfnst(int inverse, int w, int h, int color, int lfnst_idx, int tr_idx, int intra_mode, int mode_chroma,
                int cu_log2_w, int cu_log2_h, int tree_type, int luma_mode, int16_t *coeffs)
{
  cu_info_t cu;
  fill_cu_lfnst(&cu, CU_INTRA, tr_idx, lfnst_idx, lfnst_idx, intra_mode, mode_chroma, cu_log2_w, cu_log2_h);
  if (inverse) uvg_inv_lfnst(&cu, w, h, (color_t)color, (uint16_t)lfnst_idx, coeffs, (enum uvg_tree_type)tree_type, (int8_t)luma_mode);
  else uvg_fwd_lfnst(&cu, w, h, (color_t)color, (uint16_t)lfnst_idx, coeffs, (enum uvg_tree_type)tree_type, (int8_t)luma_mode);
}

int refh_quantize_residual_lfnst(const char *strategy, int cu_type, int tr_idx, int lfnst_idx, int cr_lfnst_idx,
                                 int intra_mode, int mode_chroma, int cu_log2_w, int cu_log2_h, int tree_type,
                                 int luma_mode, int w, int h, int color, int in_stride, int out_stride,
                                 const uvg_pixel *ref_in, const uvg_pixel *pred_in, uvg_pixel *rec_out,
                                 int16_t *coeff_out, int early_skip)
{
  quant_residual_func *f = refh_get("quantize_residual", strategy);
  if (!f) return -1;
  cu_info_t cu;
  fill_cu_lfnst(&cu, cu_type, tr_idx, lfnst_idx, cr_lfnst_idx, intra_mode, mode_chroma, cu_log2_w, cu_log2_h);
  g_state->collocated_luma_mode = (int8_t)luma_mode;
  return (int)f(g_state, &cu, w, h, (color_t)color, SCAN_DIAG, 0, in_stride, out_stride, ref_in, pred_in, rec_out,
                coeff_out, early_skip != 0, 0, (enum uvg_tree_type)tree_type);
}

/* uvg_image_calc_sad on a small picture pair (frame-edge semantics, image.c:438). */
unsigned refh_image_calc_sad(const uvg_pixel *pic_y, const uvg_pixel *ref_y, int pic_w,
                             int pic_h, int pic_x, int pic_y0, int ref_x, int ref_y0,
                             int bw, int bh)
{
  uvg_picture *pic = uvg_image_alloc(UVG_CSP_420, pic_w, pic_h);
  uvg_picture *ref = uvg_image_alloc(UVG_CSP_420, pic_w, pic_h);
  for (int y = 0; y < pic_h; ++y) {
    memcpy(&pic->y[y * pic->stride], &pic_y[y * pic_w], pic_w * sizeof(uvg_pixel));
    memcpy(&ref->y[y * ref->stride], &ref_y[y * pic_w], pic_w * sizeof(uvg_pixel));
  }
  unsigned r = uvg_image_calc_sad(pic, ref, pic_x, pic_y0, ref_x, ref_y0, bw, bh, NULL);
  uvg_image_free(pic);
  uvg_image_free(ref);
  return r;
}

/* encoder_control->qp_map[0] (chroma QP mapping table, encoder.c:663), or NULL. */
const int8_t *refh_chroma_qp_map(void)
{
  return g_state->encoder_control->qp_map[0];
}

/* uvg_intra_predict (intra.c:1372) on caller-supplied reference rows: top[2w+1], left[2h+1]
 * (index 0 = corner), luma, mrl 0, no ISP / MIP - what search_intra_rough calls per mode. */
void refh_intra_predict(int w, int h, int mode, const uvg_pixel *top, const uvg_pixel *left, uvg_pixel *dst)
{
  static uvg_intra_references refs;
  memset(&refs, 0, sizeof(refs));
  memcpy(refs.ref.top, top, (2 * w + 1) * sizeof(uvg_pixel));
  memcpy(refs.ref.left, left, (2 * h + 1) * sizeof(uvg_pixel));
  /* beyond the 2n+1 samples the encoder replicates the last one up to 2n + n + 2
   * (uvg_intra_build_reference_any, intra.c:879-886, :1044-1051) */
  for (int i = 2 * w + 1; i < 3 * w + 3; ++i) refs.ref.top[i] = top[2 * w];
  for (int i = 2 * h + 1; i < 3 * h + 3; ++i) refs.ref.left[i] = left[2 * h];
  cu_loc_t loc;
  uvg_cu_loc_ctor(&loc, 0, 0, w, h);
  intra_search_data_t data;
  memset(&data, 0, sizeof(data));
  data.pred_cu.type = CU_INTRA;
  data.pred_cu.log2_width = uvg_g_convert_to_log2[w];
  data.pred_cu.log2_height = uvg_g_convert_to_log2[h];
  data.pred_cu.intra.mode = (int8_t)mode;
  uvg_intra_predict(g_state, &refs, &loc, &loc, COLOR_Y, dst, &data, NULL);
}

/* The three intra strategy leaves with their own arguments (strategies-intra.h:46-80). w x h is the
 * block in the plane that is predicted: for chroma the cu_loc is built twice as large so that its
 * chroma_width / chroma_height are w x h. `above` / `left` must hold 2*128+3+33*3 samples like
 * uvg_intra_ref (intra.h:47-50). */
static void leaf_loc(cu_loc_t *loc, int w, int h, int channel_type)
{
  if (channel_type == 0) uvg_cu_loc_ctor(loc, 0, 0, w, h);
  else uvg_cu_loc_ctor(loc, 0, 0, 2 * w, 2 * h);
}
void refh_angular_pred_leaf(int w, int h, int mode, int channel_type, const uvg_pixel *above, const uvg_pixel *left,
                            uvg_pixel *dst, int mrl, int isp)
{
  cu_loc_t loc;
  leaf_loc(&loc, w, h, channel_type);
  uvg_angular_pred(&loc, (int_fast8_t)mode, (int_fast8_t)channel_type, above, left, dst, (uint8_t)mrl, (uint8_t)isp, w * h);
}
void refh_intra_pred_planar_leaf(int w, int h, int channel_type, const uvg_pixel *top, const uvg_pixel *left, uvg_pixel *dst)
{
  cu_loc_t loc;
  leaf_loc(&loc, w, h, channel_type);
  uvg_intra_pred_planar(&loc, (color_t)channel_type, top, left, dst);
}
void refh_pdpc_planar_dc_leaf(int mode, int w, int h, int channel_type, const uvg_pixel *top, const uvg_pixel *left,
                              uvg_pixel *dst)
{
  static __thread uvg_intra_ref ref;
  cu_loc_t loc;
  leaf_loc(&loc, w, h, channel_type);
  memcpy(ref.top, top, sizeof(ref.top));
  memcpy(ref.left, left, sizeof(ref.left));
  uvg_pdpc_planar_dc(mode, &loc, (color_t)channel_type, &ref, dst);
}
int refh_intra_ref_len(void) { return (int)(sizeof(((uvg_intra_ref *)0)->top) / sizeof(uvg_pixel)); }

/* strategies-ipol.h:105-108: the four MC sample-interpolation leaves. kind 0 luma, 1 luma_hi,
 * 2 chroma, 3 chroma_hi; src points at the block's top-left integer sample inside a padded window. */
void refh_sample(int kind, const uvg_pixel *src, int src_stride, int w, int h, void *dst, int dst_stride, int mvx, int mvy)
{
  const mv_t mv[2] = { mvx, mvy };
  switch (kind) {
  case 0: uvg_sample_quarterpel_luma(g_state->encoder_control, (uvg_pixel *)src, (int16_t)src_stride, w, h, (uvg_pixel *)dst, (int16_t)dst_stride, 1, 1, mv); break;
  case 1: uvg_sample_quarterpel_luma_hi(g_state->encoder_control, (uvg_pixel *)src, (int16_t)src_stride, w, h, (int16_t *)dst, (int16_t)dst_stride, 1, 1, mv); break;
  case 2: uvg_sample_octpel_chroma(g_state->encoder_control, (uvg_pixel *)src, (int16_t)src_stride, w, h, (uvg_pixel *)dst, (int16_t)dst_stride, 1, 1, mv); break;
  default: uvg_sample_octpel_chroma_hi(g_state->encoder_control, (uvg_pixel *)src, (int16_t)src_stride, w, h, (int16_t *)dst, (int16_t)dst_stride, 1, 1, mv); break;
  }
}

/* Fractional motion-search scoring the way search_frac does it (search_inter.c:1085-1250), with the
 * reference's own leaves: uvg_sample_quarterpel_luma into a block buffer, then satd_any_size against
 * the original. `ref` is the reference plane with a replicated border of `margin` samples (what
 * uvg_get_extended_block provides); bx/by are positions inside the un-padded plane. Re-entrant. */
void refh_frac_search_batch(int w, int h, const uvg_pixel *cur, int cur_stride, const uvg_pixel *ref, int ref_stride,
                            int margin, const int32_t *bx, const int32_t *by, const int32_t *mvx, const int32_t *mvy,
                            const int8_t *offs, int k, long n, uint32_t *out)
{
  uvg_pixel *pred = NULL;
  if (posix_memalign((void **)&pred, 64, 64 * 64 * sizeof(uvg_pixel))) abort();
  for (long i = 0; i < n; ++i) {
    const uvg_pixel *o = cur + (long)by[i] * cur_stride + bx[i];
    for (int c = 0; c < k; ++c) {
      const int cx = mvx[i] + offs[2 * c], cy = mvy[i] + offs[2 * c + 1];
      const mv_t mv[2] = { cx, cy };
      const uvg_pixel *src = ref + (long)(by[i] + (cy >> 4) + margin) * ref_stride + bx[i] + (cx >> 4) + margin;
      uvg_sample_quarterpel_luma(g_state->encoder_control, (uvg_pixel *)src, (int16_t)ref_stride, w, h, pred, (int16_t)w, 1, 1, mv);
      out[i * k + c] = uvg_satd_any_size(w, h, o, cur_stride, pred, w);
    }
  }
  free(pred);
}

/* crc32c_4x4 / crc32c_8x8 and bipred_average through the strategy pointers (strategies-picture.h:136-162).
 * bipred: one plane at a time through the luma path of the composite (pu at the LCU origin). */
uint32_t refh_crc32c(int n, const uvg_pixel *buf, uint32_t stride)
{
  return n == 4 ? uvg_crc32c_4x4(buf, stride) : uvg_crc32c_8x8(buf, stride);
}
void refh_bipred_average(uvg_pixel *dst /* LCU_WIDTH-strided */, const void *src0, int is_im0, const void *src1, int is_im1,
                         int w, int h)
{
  static __thread lcu_t lcu;
  yuv_t px0 = { 0 }, px1 = { 0 };
  yuv_im_t im0 = { 0 }, im1 = { 0 };
  if (is_im0) im0.y = (uvg_pixel_im *)src0; else px0.y = (uvg_pixel *)src0;
  if (is_im1) im1.y = (uvg_pixel_im *)src1; else px1.y = (uvg_pixel *)src1;
  uvg_bipred_average(&lcu, &px0, &px1, &im0, &im1, 0, 0, w, h, is_im0 ? 1 : 0, is_im1 ? 1 : 0, true, false);
  for (int y = 0; y < h; ++y) memcpy(dst + y * LCU_WIDTH, lcu.rec.y + y * LCU_WIDTH, w * sizeof(uvg_pixel));
}

/* The FME block filters the way search_frac chains them (search_inter.c:1130-1160): steps 0..step are run
 * in order on the same intermediate buffers; the four blocks of the LAST step are returned densely
 * ([4][h][w]). src = top-left of the (w+1) x (h+1) area whose sample (1,1) is the block's integer
 * position, inside a window with 4 more samples on every side. off_x / off_y: the best half-pel offset
 * (quarter-pel steps only). */
void refh_fme_filter(int step, const uvg_pixel *src, int src_stride, int w, int h, int fme_level, int off_x, int off_y,
                     uvg_pixel *out)
{
  static __thread uvg_pixel filtered[4][LCU_LUMA_SIZE];
  static __thread int16_t intermediate[5][UVG_IPOL_MAX_IM_SIZE_LUMA_SIMD];
  static __thread int16_t first_cols[5][UVG_EXT_BLOCK_W_LUMA + 1];
  ipol_blocks_func *steps[4] = { uvg_filter_hpel_blocks_hor_ver_luma, uvg_filter_hpel_blocks_diag_luma,
                                 uvg_filter_qpel_blocks_hor_ver_luma, uvg_filter_qpel_blocks_diag_luma };
  for (int st = 0; st <= step; ++st)
    steps[st](g_state->encoder_control, (uvg_pixel *)src, (int16_t)src_stride, w, h, filtered, intermediate, (int8_t)fme_level,
              first_cols, (int8_t)(st >= 2 ? off_x : 0), (int8_t)(st >= 2 ? off_y : 0));
  for (int j = 0; j < 4; ++j)
    for (int y = 0; y < h; ++y) memcpy(out + ((size_t)j * h + y) * w, &filtered[j][y * LCU_WIDTH], w * sizeof(uvg_pixel));
}

/* Rough intra search of a slice of blocks, the way search_intra_rough scores candidates
 * (search_intra.c:1055-1062 + get_cost :149-163): uvg_intra_predict for each of k modes, then
 * min(SATD, 2 * SAD) against the original block. Re-entrant (one call per worker thread).
 *   refs : [nb][2][2n+4] top / left rows as in uvg_b200_intra_rough_*;  orig : [nb][n*n] dense */
void refh_intra_rough_batch(int n, const int8_t *modes, int k, const uvg_pixel *refs_in, const uvg_pixel *orig,
                            long nb, uint32_t *costs)
{
  cost_pixel_nxn_multi_func *satd = uvg_pixels_get_satd_dual_func(n, n);
  cost_pixel_nxn_multi_func *sad = uvg_pixels_get_sad_dual_func(n, n);
  uvg_intra_references refs;
  uvg_pixel *preds = NULL;
  if (posix_memalign((void **)&preds, 64, 2 * 32 * 32 * sizeof(uvg_pixel))) abort();
  memset(preds, 0, 2 * 32 * 32 * sizeof(uvg_pixel));
  cu_loc_t loc;
  uvg_cu_loc_ctor(&loc, 0, 0, n, n);
  intra_search_data_t data;
  memset(&data, 0, sizeof(data));
  data.pred_cu.type = CU_INTRA;
  data.pred_cu.log2_width = data.pred_cu.log2_height = uvg_g_convert_to_log2[n];
  const int gp = 2 * n + 4;
  for (long b = 0; b < nb; ++b) {
    memset(&refs, 0, sizeof(refs));
    const uvg_pixel *top = refs_in + b * 2 * gp, *left = top + gp;
    memcpy(refs.ref.top, top, (2 * n + 1) * sizeof(uvg_pixel));
    memcpy(refs.ref.left, left, (2 * n + 1) * sizeof(uvg_pixel));
    for (int i = 2 * n + 1; i < 3 * n + 3; ++i) { refs.ref.top[i] = top[2 * n]; refs.ref.left[i] = left[2 * n]; }
    for (int m = 0; m < k; ++m) {
      data.pred_cu.intra.mode = modes[m];
      uvg_intra_predict(g_state, &refs, &loc, &loc, COLOR_Y, preds, &data, NULL);
      memcpy(preds + 32 * 32, preds, n * n * sizeof(uvg_pixel));
      unsigned s[2], t[2];
      satd((const uvg_pixel(*)[32 * 32])preds, orig + b * n * n, 2, s);
      sad((const uvg_pixel(*)[32 * 32])preds, orig + b * n * n, 2, t);
      costs[b * k + m] = s[0] < 2 * t[0] ? s[0] : 2 * t[0];
    }
  }
  free(preds);
}

const uint32_t *refh_scan_table(int log2_w, int log2_h)
{
  return uvg_get_scan_order_table(SCAN_GROUP_4X4, SCAN_DIAG, log2_w, log2_h, 0);
}

/* ------------------------------------------------------------------------------------
 * CPU baseline: the reference's own generic (or avx2) leaf functions timed on the host,
 * on the per-CTU all-intra hot-path workload defined in DESIGN.md ("ctu_pass").
 * Layout of the inputs mirrors uvg266_b200's batched API so both arms see the same bytes:
 *   rough search : for n in {8,16}: blocks[nb][n*n] originals, preds[nb][modes][n*n]
 *   residual     : for n in {4,8,16,32}: ref[nt][n*n], pred[nt][n*n]
 * Each worker thread takes a contiguous slice of CTUs.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int n;              /* block size */
  int count;          /* blocks per CTU */
  int modes;          /* prediction candidates per block (rough search only) */
} refh_level_t;

typedef struct {
  const char *strategy;
  int ctu_begin, ctu_end;
  int n_rough; const refh_level_t *rough;
  const uvg_pixel *const *rough_orig;   /* [level] -> [ctu][count][n*n] */
  const uvg_pixel *const *rough_pred;   /* [level] -> [ctu][count][modes][n*n] */
  uint32_t *const *rough_cost;          /* [level] -> [ctu][count][modes] */
  int n_tu; const refh_level_t *tu;
  const uvg_pixel *const *tu_ref;       /* [level] -> [ctu][count][n*n] */
  const uvg_pixel *const *tu_pred;
  uvg_pixel *const *tu_rec;
  int16_t *const *tu_coeff;
  uint32_t *const *tu_ssd;              /* [level] -> [ctu][count] */
  uint32_t *const *tu_cost;             /* [level] -> [ctu][count] (fast_coeff_cost) */
  uint64_t weights;
  int *next_unit;                       /* shared work counter */
  int pin_cpu;                          /* unused */
  int static_threads, worker_id;        /* static_threads > 0: static CTU slices */
  const struct frame_job *fj;           /* non-NULL: frame pass instead of the ctu_pass */
} refh_job_t;

/* Work is handed out dynamically: unit u = (level, chunk of CTU_CHUNK CTUs), popped from one
 * atomic counter, so a thread that is scheduled late (or shares a core) does not stall the
 * step - the same load balancing the encoder's own threadqueue gives its per-CTU jobs. */
#define CTU_CHUNK 4
static void ctu_pass_range(refh_job_t *j, int level, int c0, int c1)
{
  cu_info_t cu;
  fill_cu(&cu, CU_INTRA, 0, 0, 0, 0);
  if (level < j->n_rough) {
    const int l = level;
    const int n = j->rough[l].n, cnt = j->rough[l].count, modes = j->rough[l].modes;
    char tname[32];
    snprintf(tname, sizeof tname, "satd_%dx%d_dual", n, n);
    cost_pixel_nxn_multi_func *satd2 = refh_get(tname, j->strategy);
    if (!satd2) satd2 = refh_get(tname, "generic");
    snprintf(tname, sizeof tname, "sad_%dx%d_dual", n, n);
    cost_pixel_nxn_multi_func *sad2 = refh_get(tname, j->strategy);
    if (!sad2) sad2 = refh_get(tname, "generic");
    /* pred_buffer = uvg_pixel(*)[32*32]: the two candidates of a dual call live 1024 pixels
     * apart (strategies-picture.h:48, search_intra.c:1012); copy pairs into such a buffer
     * like the encoder's rough search does when it predicts into preds[0..1]. */
    ALIGNED(64) uvg_pixel pb[2][32 * 32];
    for (int c = c0; c < c1; ++c) {
      for (int b = 0; b < cnt; ++b) {
        const size_t bi = (size_t)c * cnt + b;
        const uvg_pixel *orig = j->rough_orig[l] + bi * n * n;
        for (int m = 0; m < modes; m += 2) {
          const uvg_pixel *p0 = j->rough_pred[l] + (bi * modes + m) * n * n;
          const int m1 = (m + 1 < modes) ? m + 1 : m;
          const uvg_pixel *p1 = j->rough_pred[l] + (bi * modes + m1) * n * n;
          memcpy(pb[0], p0, n * n * sizeof(uvg_pixel));
          memcpy(pb[1], p1, n * n * sizeof(uvg_pixel));
          unsigned satd[2], sad[2];
          satd2((const pred_buffer)pb, orig, 2, satd);
          sad2((const pred_buffer)pb, orig, 2, sad);
          /* get_cost_dual, search_intra.c:133-192: cost = MIN(satd, 2*sad) */
          uint32_t *out = j->rough_cost[l] + bi * modes;
          out[m] = MIN(satd[0], 2 * sad[0]);
          if (m1 != m) out[m1] = MIN(satd[1], 2 * sad[1]);
        }
      }
    }
    return;
  }
  const int l = level - j->n_rough;
  quant_residual_func *qr = refh_get("quantize_residual", j->strategy);
  pixels_calc_ssd_func *ssd = refh_get("pixels_calc_ssd", j->strategy);
  fast_coeff_cost_func *fcc = refh_get("fast_coeff_cost", j->strategy);
  if (!fcc) fcc = refh_get("fast_coeff_cost", "generic");
  if (!qr) qr = refh_get("quantize_residual", "generic");
  if (!ssd) ssd = refh_get("pixels_calc_ssd", "generic");
  const int n = j->tu[l].n, cnt = j->tu[l].count;
  for (int c = c0; c < c1; ++c) {
    for (int b = 0; b < cnt; ++b) {
      const size_t bi = (size_t)c * cnt + b;
      const uvg_pixel *ref = j->tu_ref[l] + bi * n * n;
      const uvg_pixel *pred = j->tu_pred[l] + bi * n * n;
      uvg_pixel *rec = j->tu_rec[l] + bi * n * n;
      int16_t *coeff = j->tu_coeff[l] + bi * n * n;
      qr(g_state, &cu, n, n, COLOR_Y, SCAN_DIAG, 0, n, n, ref, pred, rec, coeff, false, 0,
         UVG_BOTH_T);
      j->tu_ssd[l][bi] = ssd(ref, rec, n, n, n, n);
      j->tu_cost[l][bi] = fcc(coeff, n, n, j->weights);
    }
  }
}

/* ------------------------------------------------------------------------------------
 * Frame pass: the CPU arm of the device-resident frame pipeline (uvg_b200_frame_*, include/uvg266_b200.h
 * section 4), run with the reference's own functions - the same per-CTU work, the same open-loop reference
 * samples (source-picture neighbours, border replication), the same CTU-major output layout:
 *   rough search  : uvg_intra_predict for the level's candidate modes, two at a time into a pred_buffer like
 *                   search_intra_rough (search_intra.c:1012-1062), cost = MIN(satd, 2 * sad) from the dual functions
 *   selection     : the n_best cheapest candidates, ties in candidate order
 *   residual pass : uvg_intra_predict with the TU's mode -> quantize_residual -> pixels_calc_ssd -> fast_coeff_cost
 * ---------------------------------------------------------------------------------- */
typedef struct { int n, k; int8_t modes[68]; } refh_frame_level_t;

typedef struct frame_job {
  const uvg_pixel *plane; int W, H, stride, ctu_cols, ctu_begin, ctu_end;
  int n_levels; const refh_frame_level_t *levels;
  uint32_t *const *rough_cost; uint8_t *const *rough_best;
  int n_tus; const refh_level_t *tus; int n_best;
  uint8_t *const *tu_mode; uint8_t *const *tu_has; uint32_t *const *tu_ssd; uint32_t *const *tu_cost;
  int16_t *const *tu_coeff; uvg_pixel *const *tu_rec;
  uint64_t weights;
} frame_job_t;

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

static void frame_refs(const frame_job_t *f, int X, int Y, int n, uvg_intra_references *refs)
{
  memset(refs, 0, sizeof(*refs));
  for (int i = 0; i < 3 * n + 3; ++i) {
    const int k = i < 2 * n ? i : 2 * n;              /* the rows continue with the last sample, intra.c:879-886 */
    refs->ref.top[i] = f->plane[(size_t)clampi(Y - 1, 0, f->H - 1) * f->stride + clampi(X - 1 + k, 0, f->W - 1)];
    refs->ref.left[i] = f->plane[(size_t)clampi(Y - 1 + k, 0, f->H - 1) * f->stride + clampi(X - 1, 0, f->W - 1)];
  }
}

static void frame_block(const frame_job_t *f, int X, int Y, int n, uvg_pixel *dst)
{
  for (int y = 0; y < n; ++y)
    for (int x = 0; x < n; ++x)
      dst[y * n + x] = f->plane[(size_t)MIN(Y + y, f->H - 1) * f->stride + MIN(X + x, f->W - 1)];
}

static char g_frame_strategy_buf[32];
static const char *g_frame_strategy = "generic";   /* refh_frame_strategy() */

static void frame_ctu(const frame_job_t *f, int ctu)
{
  const int cx = (ctu % f->ctu_cols) * LCU_WIDTH, cy = (ctu / f->ctu_cols) * LCU_WIDTH;
  const int rel = ctu - f->ctu_begin;
  uvg_intra_references refs;
  ALIGNED(64) uvg_pixel preds[2][32 * 32];
  ALIGNED(64) uvg_pixel orig[32 * 32];
  intra_search_data_t data;
  memset(&data, 0, sizeof(data));
  data.pred_cu.type = CU_INTRA;
  for (int l = 0; l < f->n_levels; ++l) {
    const int n = f->levels[l].n, k = f->levels[l].k, bpr = LCU_WIDTH / n, ppc = bpr * bpr;
    cost_pixel_nxn_multi_func *satd = uvg_pixels_get_satd_dual_func(n, n);
    cost_pixel_nxn_multi_func *sad = uvg_pixels_get_sad_dual_func(n, n);
    cu_loc_t loc;
    uvg_cu_loc_ctor(&loc, 0, 0, n, n);
    data.pred_cu.log2_width = data.pred_cu.log2_height = uvg_g_convert_to_log2[n];
    for (int j = 0; j < ppc; ++j) {
      const int X = cx + (j % bpr) * n, Y = cy + (j / bpr) * n;
      const size_t b = (size_t)rel * ppc + j;
      frame_refs(f, X, Y, n, &refs);
      frame_block(f, X, Y, n, orig);
      uint32_t *cost = f->rough_cost[l] + b * k;
      for (int m = 0; m < k; m += 2) {
        const int m1 = m + 1 < k ? m + 1 : m;
        data.pred_cu.intra.mode = f->levels[l].modes[m];
        uvg_intra_predict(g_state, &refs, &loc, &loc, COLOR_Y, preds[0], &data, NULL);
        data.pred_cu.intra.mode = f->levels[l].modes[m1];
        uvg_intra_predict(g_state, &refs, &loc, &loc, COLOR_Y, preds[1], &data, NULL);
        unsigned s[2], t[2];
        satd((const pred_buffer)preds, orig, 2, s);
        sad((const pred_buffer)preds, orig, 2, t);
        cost[m] = MIN(s[0], 2 * t[0]);
        if (m1 != m) cost[m1] = MIN(s[1], 2 * t[1]);
      }
      /* the n_best cheapest, ties in candidate order */
      uint8_t *best = f->rough_best[l] + b * f->n_best;
      uint32_t bc[4] = { UINT32_MAX, UINT32_MAX, UINT32_MAX, UINT32_MAX };
      int bi[4] = { 0, 0, 0, 0 };
      for (int m = 0; m < k; ++m) {
        uint32_t c = cost[m];
        int idx = m;
        for (int s2 = 0; s2 < f->n_best; ++s2)
          if (c < bc[s2]) { const uint32_t tc = bc[s2]; const int ti = bi[s2]; bc[s2] = c; bi[s2] = idx; c = tc; idx = ti; }
      }
      for (int s2 = 0; s2 < f->n_best; ++s2) best[s2] = (uint8_t)f->levels[l].modes[bi[s2 < k ? s2 : k - 1]];
    }
  }
  quant_residual_func *qr = refh_get("quantize_residual", g_frame_strategy);
  pixels_calc_ssd_func *ssd = refh_get("pixels_calc_ssd", g_frame_strategy);
  fast_coeff_cost_func *fcc = refh_get("fast_coeff_cost", g_frame_strategy);
  if (!qr) qr = refh_get("quantize_residual", "generic");
  if (!ssd) ssd = refh_get("pixels_calc_ssd", "generic");
  if (!fcc) fcc = refh_get("fast_coeff_cost", "generic");
  cu_info_t cu;
  fill_cu(&cu, CU_INTRA, 0, 0, 0, 0);
  for (int i = 0; i < f->n_tus; ++i) {
    const int n = f->tus[i].n, per_ctu = f->tus[i].count, bpr = LCU_WIDTH / n, ppc = bpr * bpr;
    int src = -1;                                      /* mode source level: smallest level >= n, else the largest */
    for (int l = 0; l < f->n_levels; ++l) if (f->levels[l].n >= n && (src < 0 || f->levels[l].n < f->levels[src].n)) src = l;
    if (src < 0) for (int l = 0; l < f->n_levels; ++l) if (src < 0 || f->levels[l].n > f->levels[src].n) src = l;
    const int sn = f->levels[src].n, sb = LCU_WIDTH / sn;
    cu_loc_t loc;
    uvg_cu_loc_ctor(&loc, 0, 0, n, n);
    data.pred_cu.log2_width = data.pred_cu.log2_height = uvg_g_convert_to_log2[n];
    for (int j = 0; j < per_ctu; ++j) {
      const int p = j % ppc, rank = MIN(j / ppc, f->n_best - 1);
      const int X = cx + (p % bpr) * n, Y = cy + (p / bpr) * n;
      const size_t t = (size_t)rel * per_ctu + j;
      const size_t sblk = (size_t)rel * sb * sb + ((Y - cy) / sn) * sb + (X - cx) / sn;
      const int mode = f->rough_best[src][sblk * f->n_best + rank];
      frame_refs(f, X, Y, n, &refs);
      frame_block(f, X, Y, n, orig);
      data.pred_cu.intra.mode = (int8_t)mode;
      uvg_intra_predict(g_state, &refs, &loc, &loc, COLOR_Y, preds[0], &data, NULL);
      uvg_pixel *rec = f->tu_rec[i] + t * n * n;
      int16_t *coeff = f->tu_coeff[i] + t * n * n;
      f->tu_mode[i][t] = (uint8_t)mode;
      f->tu_has[i][t] = (uint8_t)qr(g_state, &cu, n, n, COLOR_Y, SCAN_DIAG, 0, n, n, orig, preds[0], rec, coeff, false, 0, UVG_BOTH_T);
      f->tu_ssd[i][t] = ssd(orig, rec, n, n, n, n);
      f->tu_cost[i][t] = fcc(coeff, n, n, f->weights);
    }
  }
}

static void *ctu_pass_worker(void *arg)
{
  refh_job_t *j = arg;
  if (j->fj) {                                         /* frame pass: one unit = one CTU */
    const int n_units = j->fj->ctu_end - j->fj->ctu_begin;
    for (;;) {
      const int u = __atomic_fetch_add(j->next_unit, 1, __ATOMIC_RELAXED);
      if (u >= n_units) break;
      frame_ctu(j->fj, j->fj->ctu_begin + u);
    }
    return NULL;
  }
  const int n_ctu = j->ctu_end;
  if (j->static_threads > 0) {
    /* static partition: worker w always owns the same contiguous CTU slice (NUMA-friendly:
     * with pinned workers the pages a slice touches stay on that worker's node) */
    const int w = j->worker_id, nw = j->static_threads;
    const int c0 = (int)((long long)n_ctu * w / nw), c1 = (int)((long long)n_ctu * (w + 1) / nw);
    for (int level = 0; level < j->n_rough + j->n_tu; ++level) ctu_pass_range(j, level, c0, c1);
    return NULL;
  }
  const int chunks = (n_ctu + CTU_CHUNK - 1) / CTU_CHUNK;
  const int n_units = chunks * (j->n_rough + j->n_tu);
  for (;;) {
    const int u = __atomic_fetch_add(j->next_unit, 1, __ATOMIC_RELAXED);
    if (u >= n_units) break;
    const int level = u / chunks, ch = u % chunks;
    const int c0 = ch * CTU_CHUNK, c1 = MIN(n_ctu, c0 + CTU_CHUNK);
    ctu_pass_range(j, level, c0, c1);
  }
  return NULL;
}

/* Persistent worker pool (like the encoder's own threadqueue, threadqueue.c:275): threads are
 * created once and parked on a condition variable; creating 128 fresh threads per pass made the
 * pass time alternate 5x between consecutive calls on the 2-socket host. */
#define POOL_MAX 512
static struct {
  pthread_t th[POOL_MAX];
  int n_threads;                 /* created so far */
  pthread_mutex_t mu;
  pthread_cond_t cv_work, cv_done;
  unsigned long generation;      /* bumped per job */
  int active;                    /* workers taking part in the current job */
  int remaining;                 /* workers still running the current job */
  refh_job_t job;
} g_pool = { .mu = PTHREAD_MUTEX_INITIALIZER, .cv_work = PTHREAD_COND_INITIALIZER, .cv_done = PTHREAD_COND_INITIALIZER };

static void *pool_main(void *arg)
{
  const int id = (int)(intptr_t)arg;
  unsigned long seen = 0;
  const char *pin_env = getenv("REFH_PIN");
  if (pin_env && pin_env[0] == '1') {
    cpu_set_t set;
    CPU_ZERO(&set); CPU_SET(id % (int)sysconf(_SC_NPROCESSORS_ONLN), &set);
    pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
  }
  for (;;) {
    pthread_mutex_lock(&g_pool.mu);
    while (g_pool.generation == seen) pthread_cond_wait(&g_pool.cv_work, &g_pool.mu);
    seen = g_pool.generation;
    const int take_part = id < g_pool.active;
    refh_job_t job = g_pool.job;
    pthread_mutex_unlock(&g_pool.mu);
    if (!take_part) continue;
    job.pin_cpu = -1;
    job.worker_id = id;
    ctu_pass_worker(&job);
    pthread_mutex_lock(&g_pool.mu);
    if (--g_pool.remaining == 0) pthread_cond_signal(&g_pool.cv_done);
    pthread_mutex_unlock(&g_pool.mu);
  }
  return NULL;
}

/* Runs the ctu_pass over `n_ctu` CTUs on `threads` pool workers; returns seconds. */
double refh_ctu_pass(const char *strategy, int threads, int n_ctu,
                     int n_rough, const refh_level_t *rough,
                     const uvg_pixel *const *rough_orig, const uvg_pixel *const *rough_pred,
                     uint32_t *const *rough_cost,
                     int n_tu, const refh_level_t *tu,
                     const uvg_pixel *const *tu_ref, const uvg_pixel *const *tu_pred,
                     uvg_pixel *const *tu_rec, int16_t *const *tu_coeff,
                     uint32_t *const *tu_ssd, uint32_t *const *tu_cost, uint64_t weights)
{
  if (threads < 1) threads = 1;
  if (threads > POOL_MAX) threads = POOL_MAX;
  pthread_mutex_lock(&g_pool.mu);
  while (g_pool.n_threads < threads) {
    pthread_create(&g_pool.th[g_pool.n_threads], NULL, pool_main, (void *)(intptr_t)g_pool.n_threads);
    ++g_pool.n_threads;
  }
  pthread_mutex_unlock(&g_pool.mu);

  struct timespec t0, t1;
  int next_unit = 0;
  refh_job_t j;
  memset(&j, 0, sizeof j);
  j.strategy = strategy;
  j.ctu_begin = 0; j.ctu_end = n_ctu;
  j.next_unit = &next_unit;
  j.n_rough = n_rough; j.rough = rough;
  j.rough_orig = rough_orig; j.rough_pred = rough_pred; j.rough_cost = rough_cost;
  j.n_tu = n_tu; j.tu = tu;
  j.tu_ref = tu_ref; j.tu_pred = tu_pred; j.tu_rec = tu_rec; j.tu_coeff = tu_coeff;
  j.tu_ssd = tu_ssd; j.tu_cost = tu_cost; j.weights = weights;
  {
    const char *st = getenv("REFH_STATIC");
    j.static_threads = (st && st[0] == '1') ? threads : 0;
  }
  clock_gettime(CLOCK_MONOTONIC, &t0);
  pthread_mutex_lock(&g_pool.mu);
  g_pool.job = j;
  g_pool.active = threads;
  g_pool.remaining = threads;
  ++g_pool.generation;
  pthread_cond_broadcast(&g_pool.cv_work);
  while (g_pool.remaining > 0) pthread_cond_wait(&g_pool.cv_done, &g_pool.mu);
  pthread_mutex_unlock(&g_pool.mu);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

static double run_pool_job(refh_job_t *j, int threads)
{
  if (threads < 1) threads = 1;
  if (threads > POOL_MAX) threads = POOL_MAX;
  pthread_mutex_lock(&g_pool.mu);
  while (g_pool.n_threads < threads) {
    pthread_create(&g_pool.th[g_pool.n_threads], NULL, pool_main, (void *)(intptr_t)g_pool.n_threads);
    ++g_pool.n_threads;
  }
  pthread_mutex_unlock(&g_pool.mu);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  pthread_mutex_lock(&g_pool.mu);
  g_pool.job = *j;
  g_pool.active = threads;
  g_pool.remaining = threads;
  ++g_pool.generation;
  pthread_cond_broadcast(&g_pool.cv_work);
  while (g_pool.remaining > 0) pthread_cond_wait(&g_pool.cv_done, &g_pool.mu);
  pthread_mutex_unlock(&g_pool.mu);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* Which registered implementations the frame pass calls: "generic" (the parity oracle, default) or "avx2" (the
 * reference's own SIMD strategies, for context; names without an avx2 entry stay generic). Also repoints the globals
 * uvg_intra_predict / uvg_pixels_get_* dispatch through. Returns how many globals now point at `strategy`. */
int refh_frame_strategy(const char *strategy)
{
  refh_use("generic");
  g_frame_strategy = "generic";
  if (strategy && strcmp(strategy, "generic")) {
    const int n = refh_use(strategy);
    snprintf(g_frame_strategy_buf, sizeof g_frame_strategy_buf, "%s", strategy);
    g_frame_strategy = g_frame_strategy_buf;
    return n;
  }
  return 0;
}

/* Runs the frame pass over CTUs [ctu_begin, ctu_end) on `threads` pool workers; returns seconds. */
double refh_frame_pass(int threads, const uvg_pixel *plane, int W, int H, int stride, int ctu_begin, int ctu_end,
                       int n_levels, const refh_frame_level_t *levels, uint32_t *const *rough_cost,
                       uint8_t *const *rough_best, int n_tus, const refh_level_t *tus, int n_best,
                       uint8_t *const *tu_mode, uint8_t *const *tu_has, uint32_t *const *tu_ssd,
                       uint32_t *const *tu_cost, int16_t *const *tu_coeff, uvg_pixel *const *tu_rec, uint64_t weights)
{
  frame_job_t f = { plane, W, H, stride, (W + LCU_WIDTH - 1) / LCU_WIDTH, ctu_begin, ctu_end, n_levels, levels, rough_cost,
                    rough_best, n_tus, tus, n_best, tu_mode, tu_has, tu_ssd, tu_cost, tu_coeff, tu_rec, weights };
  int next_unit = 0;
  refh_job_t j;
  memset(&j, 0, sizeof j);
  j.next_unit = &next_unit;
  j.fj = &f;
  return run_pool_job(&j, threads);
}

/* Pair-batch micro-baseline: n_pairs dense n*n block pairs through one nxn cost function
 * ("sad_64x64", "satd_64x64", ...), split over threads; returns seconds. */
typedef struct {
  cost_pixel_nxn_func *f; const uvg_pixel *a, *b; uint32_t *out; size_t begin, end; int nn;
} pair_job_t;

static void *pair_worker(void *arg)
{
  pair_job_t *j = arg;
  for (size_t i = j->begin; i < j->end; ++i)
    j->out[i] = j->f(j->a + i * j->nn, j->b + i * j->nn);
  return NULL;
}

double refh_pair_batch(const char *type, const char *strategy, int threads, int n,
                       size_t n_pairs, const uvg_pixel *a, const uvg_pixel *b, uint32_t *out)
{
  cost_pixel_nxn_func *f = refh_get(type, strategy);
  if (!f) return -1.0;
  if (threads < 1) threads = 1;
  pthread_t *th = calloc(threads, sizeof(*th));
  pair_job_t *jobs = calloc(threads, sizeof(*jobs));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < threads; ++t) {
    jobs[t] = (pair_job_t){ f, a, b, out, n_pairs * t / threads, n_pairs * (t + 1) / threads, n * n };
    pthread_create(&th[t], NULL, pair_worker, &jobs[t]);
  }
  for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(th); free(jobs);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
