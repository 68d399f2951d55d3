/*
 * strategies_cuda.c - registers libuvg266_b200's kernels as the "cuda" strategy of
 * uvg266's strategyselector (priority 50 > avx2's 40), so the encoder's unmodified search /
 * RDO code calls them through its function-pointer table.
 *
 * Host language = the reference's (C). Compiled against the reference headers of the same
 * commit (encoder_state_t / cu_info_t are not ABI-stable); everything below the boundary is
 * reached through the plain C ABI of include/uvg266_b200.h.
 *
 * Mechanism restated from:
 *   strategyselector.h:98-99, strategyselector.c:242   uvg_strategyselector_register
 *   strategies-picture.c:92-114                         group registrars, hardware-flag guards
 *   generic/picture-generic.c:1445-1498, dct-generic.c:2681-2704, quant-generic.c:707-719
 *                                                       the type strings registered here
 * A registrar returns 0 when the device cannot be initialised: uvg_strategyselector_init
 * then fails and uvg266_open reports "Failed to initialize strategies." (uvg266.c:96-99) -
 * there is no silent CPU fallback.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "global.h"
#include "uvg266.h"
#include "cu.h"
#include "encoder.h"
#include "encoderstate.h"
#include "strategyselector.h"
#include "strategies/strategies-picture.h"
#include "strategies/strategies-dct.h"
#include "strategies/strategies-quant.h"
#include "strategies/strategies-intra.h"
#include "strategies/strategies-ipol.h"
#include "transform.h"
#include "rdo.h"
#include "dep_quant.h"
#include "reshape.h"

#include "uvg266_b200.h"

#define CUDA_PRIORITY 50
#define STRATEGY_NAME "cuda"

#if UVG_BIT_DEPTH == 8
#define B200(name) uvg_b200_##name##_8
#else
#define B200(name) uvg_b200_##name##_16
#endif

/* ---- picture --------------------------------------------------------------------- */
static unsigned reg_sad_cuda(const uvg_pixel *const d1, const uvg_pixel *const d2, const int w,
                             const int h, const unsigned s1, const unsigned s2)
{ return B200(reg_sad)(d1, d2, w, h, s1, s2); }

#define NXN(n)                                                                                   \
  static unsigned sad_##n##_cuda(const uvg_pixel *const a, const uvg_pixel *const b)             \
  { return B200(sad_##n##x##n)(a, b); }                                                          \
  static unsigned satd_##n##_cuda(const uvg_pixel *const a, const uvg_pixel *const b)            \
  { return B200(satd_##n##x##n)(a, b); }
NXN(4) NXN(8) NXN(16) NXN(32) NXN(64)

#define DUAL(n)                                                                                  \
  static void sad_##n##_dual_cuda(const pred_buffer preds, const uvg_pixel *const orig,          \
                                  unsigned num_modes, unsigned *costs_out)                       \
  { B200(sad_##n##x##n##_dual)(&preds[0][0], orig, num_modes, costs_out); }                      \
  static void satd_##n##_dual_cuda(const pred_buffer preds, const uvg_pixel *const orig,         \
                                   unsigned num_modes, unsigned *costs_out)                      \
  { B200(satd_##n##x##n##_dual)(&preds[0][0], orig, num_modes, costs_out); }
DUAL(4) DUAL(8) DUAL(16) DUAL(32) DUAL(64)

static unsigned satd_any_size_cuda(int w, int h, const uvg_pixel *b1, int s1, const uvg_pixel *b2, int s2)
{ return B200(satd_any_size)(w, h, b1, s1, b2, s2); }
static unsigned satd_any_size_vtm_cuda(int w, int h, const uvg_pixel *b1, int s1, const uvg_pixel *b2, int s2)
{ return B200(satd_any_size_vtm)(w, h, b1, s1, b2, s2); }
static void satd_any_size_quad_cuda(int w, int h, const uvg_pixel **preds, const int stride,
                                    const uvg_pixel *orig, const int orig_stride, unsigned num_modes,
                                    unsigned *costs_out, int8_t *valid)
{ B200(satd_any_size_quad)(w, h, preds, stride, orig, orig_stride, num_modes, costs_out, valid); }
static unsigned pixels_calc_ssd_cuda(const uvg_pixel *const ref, const uvg_pixel *const rec,
                                     const int rs, const int cs, const int w, const int h)
{ return B200(pixels_calc_ssd)(ref, rec, rs, cs, w, h); }
static optimized_sad_func_ptr_t get_optimized_sad_cuda(int32_t width)
{ return (optimized_sad_func_ptr_t)B200(get_optimized_sad)(width); }
static uint32_t ver_sad_cuda(const uvg_pixel *pic, const uvg_pixel *ref, int32_t w, int32_t h, uint32_t ps)
{ return B200(ver_sad)(pic, ref, w, h, ps); }
static uint32_t hor_sad_cuda(const uvg_pixel *pic, const uvg_pixel *ref, int32_t w, int32_t h,
                             uint32_t ps, uint32_t rs, uint32_t left, uint32_t right)
{ return B200(hor_sad)(pic, ref, w, h, ps, rs, left, right); }
static void generate_residual_cuda(const uvg_pixel *ref, const uvg_pixel *pred, int16_t *residual,
                                   int w, int h, int rs, int ps)
{ B200(generate_residual)(ref, pred, residual, w, h, rs, ps); }

/* ---- dct ------------------------------------------------------------------------- */
/* dct_func is typed `unsigned (...)` (strategies-dct.h:44); the value is never read */
#define DCTFN(name)                                                                              \
  static unsigned name##_cuda(int8_t bitdepth, const int16_t *in, int16_t *out)                  \
  { uvg_b200_##name(bitdepth, in, out); return 0; }
DCTFN(dct_4x4) DCTFN(dct_8x8) DCTFN(dct_16x16) DCTFN(dct_32x32)
DCTFN(idct_4x4) DCTFN(idct_8x8) DCTFN(idct_16x16) DCTFN(idct_32x32)
DCTFN(fast_forward_dst_4x4) DCTFN(fast_inverse_dst_4x4)

static void mts_dct_cuda(const int8_t bitdepth, const color_t color, const cu_info_t *tu,
                         const int8_t w, const int8_t h, const int16_t *in, int16_t *out,
                         const int8_t mts_type)
{
  uvg_b200_mts_dct(bitdepth, (int)color, tu->type, tu->tr_idx, tu->lfnst_idx, tu->cr_lfnst_idx,
                   tu->type == CU_INTRA ? tu->intra.isp_mode : 0, w, h, in, out, mts_type);
}
static void mts_idct_cuda(const int8_t bitdepth, const color_t color, const cu_info_t *tu,
                          const int8_t w, const int8_t h, const int16_t *in, int16_t *out,
                          const int8_t mts_type)
{
  uvg_b200_mts_idct(bitdepth, (int)color, tu->type, tu->tr_idx, tu->lfnst_idx, tu->cr_lfnst_idx,
                    tu->type == CU_INTRA ? tu->intra.isp_mode : 0, w, h, in, out, mts_type);
}

/* bipred_average (picture-generic.c:1195-1247): the composite only chooses, per plane, which of the two
 * predictions are pixels and which are 14-bit intermediates; the averaging itself is the leaf */
static void bipred_average_cuda(lcu_t *const lcu, const yuv_t *const px_L0, const yuv_t *const px_L1,
                                const yuv_im_t *const im_L0, const yuv_im_t *const im_L1, const unsigned pu_x,
                                const unsigned pu_y, const unsigned pu_w, const unsigned pu_h, const unsigned im_flags_L0,
                                const unsigned im_flags_L1, const bool predict_luma, const bool predict_chroma)
{
  if (predict_luma) {
    const unsigned off = SUB_SCU(pu_y) * LCU_WIDTH + SUB_SCU(pu_x);
    const int i0 = im_flags_L0 & 1, i1 = im_flags_L1 & 1;
    B200(bipred_average)(lcu->rec.y + off, LCU_WIDTH, i0 ? (const void *)im_L0->y : (const void *)px_L0->y, i0,
                         i1 ? (const void *)im_L1->y : (const void *)px_L1->y, i1, pu_w, pu_h);
  }
  if (predict_chroma) {
    const unsigned off = SUB_SCU(pu_y) / 2 * LCU_WIDTH_C + SUB_SCU(pu_x) / 2;
    const int i0 = (im_flags_L0 & 2) != 0, i1 = (im_flags_L1 & 2) != 0;
    B200(bipred_average)(lcu->rec.u + off, LCU_WIDTH_C, i0 ? (const void *)im_L0->u : (const void *)px_L0->u, i0,
                         i1 ? (const void *)im_L1->u : (const void *)px_L1->u, i1, pu_w / 2, pu_h / 2);
    B200(bipred_average)(lcu->rec.v + off, LCU_WIDTH_C, i0 ? (const void *)im_L0->v : (const void *)px_L0->v, i0,
                         i1 ? (const void *)im_L1->v : (const void *)px_L1->v, i1, pu_w / 2, pu_h / 2);
  }
}
static double pixel_var_cuda(const uvg_pixel *buf, const uint32_t len) { return B200(pixel_var)(buf, len); }
static uint32_t crc32c_4x4_cuda(const uvg_pixel *buf, uint32_t pic_stride) { return B200(crc32c_4x4)(buf, pic_stride); }
static uint32_t crc32c_8x8_cuda(const uvg_pixel *buf, uint32_t pic_stride) { return B200(crc32c_8x8)(buf, pic_stride); }

/* ---- quant ----------------------------------------------------------------------- */
/* The implementations the reference itself registered for the quant group (highest priority below ours),
 * captured at registration. Calls the cuda strategy does not cover are FORWARDED to them - the
 * composites (quantize_residual, quant_cbcr_residual) then run their transform / quant / dequant steps
 * through the strategy table, i.e. on the cuda leaves, and the encoder-side tools (LFNST, RDOQ, dep-quant,
 * LMCS: transform.c, rdo.c, dep_quant.c, reshape.c) run where they live. Nothing aborts mid-encode. */
static int lfnst_on_device;              /* the encoder's LFNST kernels were found and copied to the device */
static quant_func *ref_quant;
static dequant_func *ref_dequant;
static quant_residual_func *ref_quantize_residual;
static quant_cbcr_func *ref_quant_cbcr_residual;

static int scaling_lists_on(const encoder_control_t *const enc)
{
  return enc->cfg.scaling_list != UVG_SCALING_LIST_OFF || enc->scaling_list.enable;
}

/* Flatten what uvg_quant_generic / uvg_dequant_generic read through encoder_state_t
 * (quant-generic.c:62-84, 620-637). */
static void fill_cfg(const encoder_state_t *const state, const cu_info_t *const cu, uvg_b200_quant_cfg *c)
{
  const encoder_control_t *const enc = state->encoder_control;
  memset(c, 0, sizeof(*c));
  c->qp = state->qp;
  c->bitdepth = enc->bitdepth;
  c->slice_intra = state->frame->slicetype == UVG_SLICE_I;
  c->signhide = enc->cfg.signhide_enable;
  c->mts = enc->cfg.mts;
  c->isp_mode = (cu && cu->type == CU_INTRA) ? cu->intra.isp_mode : 0;
  if (enc->qp_map[0]) {
    for (int i = 0; i < 64; ++i) c->chroma_qp_map[i] = enc->qp_map[0][i];
    c->has_chroma_qp_map = 1;
  }
}

/* quant / dequant with scaling lists (quant-generic.c:74-94, :639-659) read per-position tables owned by
 * the encoder (scalinglist.c); that branch stays with the reference's implementation. */
static unsigned quant_cuda(const encoder_state_t *const state, coeff_t *coef, coeff_t *q_coef, int32_t w,
                           int32_t h, color_t color, int8_t scan_idx, int8_t block_type,
                           int8_t transform_skip, uint8_t lfnst_idx)
{
  if (scaling_lists_on(state->encoder_control))
    return ref_quant(state, coef, q_coef, w, h, color, scan_idx, block_type, transform_skip, lfnst_idx);
  uvg_b200_quant_cfg c;
  fill_cfg(state, NULL, &c);
  uvg_b200_quant(&c, coef, q_coef, w, h, (int)color, scan_idx, block_type, transform_skip, lfnst_idx);
  return 0;
}

static unsigned dequant_cuda(const encoder_state_t *const state, coeff_t *q_coef, coeff_t *coef, int32_t w,
                             int32_t h, color_t color, int8_t block_type, int8_t transform_skip)
{
  const encoder_control_t *const enc = state->encoder_control;
  /* trellis-coded quantisation has its own dequantiser in the encoder proper (dep_quant.c:1069), which
   * the reference's dequant forwards to itself (:621-624) */
  if ((enc->cfg.dep_quant && !transform_skip) || scaling_lists_on(enc))
    return ref_dequant(state, q_coef, coef, w, h, color, block_type, transform_skip);
  uvg_b200_quant_cfg c;
  fill_cfg(state, NULL, &c);
  uvg_b200_dequant(&c, q_coef, coef, w, h, (int)color, block_type, transform_skip);
  return 0;
}

/* quantize_residual: the plain chain (residual -> T -> quant -> dequant -> IT -> recon) is ONE fused
 * kernel. With a tool of the encoder proper switched on the call is forwarded (see above). */
static unsigned quantize_residual_cuda(encoder_state_t *const state, const cu_info_t *const cur_cu,
                                       const int width, const int height, const color_t color,
                                       const coeff_scan_order_t scan_order, const int use_trskip,
                                       const int in_stride, const int out_stride,
                                       const uvg_pixel *const ref_in, const uvg_pixel *const pred_in,
                                       uvg_pixel *rec_out, coeff_t *coeff_out, bool early_skip,
                                       int lmcs_chroma_adj, enum uvg_tree_type tree_type)
{
  const encoder_control_t *const enc = state->encoder_control;
  const bool lmcs = state->tile->frame->lmcs_aps->m_sliceReshapeInfo.enableChromaAdj && color != COLOR_Y;
  const bool lfnst = enc->cfg.lfnst && cur_cu->type == CU_INTRA;
  const bool rdoq = enc->cfg.rdoq_enable && ((width > 4 || !enc->cfg.rdoq_skip) || use_trskip);
  const bool dq = !use_trskip && enc->cfg.dep_quant;
  /* LFNST stays inside the fused kernel when the encoder's kernel tables could be handed to the device */
  if (lmcs || rdoq || dq || scaling_lists_on(enc) || (lfnst && (!lfnst_on_device || use_trskip)))
    return ref_quantize_residual(state, cur_cu, width, height, color, scan_order, use_trskip, in_stride, out_stride,
                                 ref_in, pred_in, rec_out, coeff_out, early_skip, lmcs_chroma_adj, tree_type);
  uvg_b200_quant_cfg c;
  fill_cfg(state, cur_cu, &c);
  if (lfnst) {
    /* what uvg_quantize_residual_generic (:503-509) and uvg_fwd_lfnst (transform.c:979-1014) derive from the CU */
    const uint8_t lfnst_index = (tree_type != UVG_CHROMA_T || color == COLOR_Y) ? cur_cu->lfnst_idx : cur_cu->cr_lfnst_idx;
    const bool sep_tree = cur_cu->log2_height + cur_cu->log2_width < 6 || tree_type != UVG_BOTH_T;
    const bool mts_skip = cur_cu->tr_idx == MTS_SKIP && color == COLOR_Y;
    int mode = color == COLOR_Y ? cur_cu->intra.mode : cur_cu->intra.mode_chroma;
    if (mode >= 81 && mode <= 83) mode = state->collocated_luma_mode;          /* CCLM */
    if (color == COLOR_Y && cur_cu->intra.mip_flag) mode = 0;                 /* MIP -> planar */
    c.cu_lfnst_idx = cur_cu->lfnst_idx;
    c.cu_cr_lfnst_idx = cur_cu->cr_lfnst_idx;
    c.lfnst_index = lfnst_index;
    c.lfnst_apply = lfnst_index && !mts_skip && (color == COLOR_Y || sep_tree);
    c.lfnst_mode = mode;
    c.lfnst_wa_log2_w = color == COLOR_Y ? cur_cu->log2_width : uvg_g_convert_to_log2[width];
    c.lfnst_wa_log2_h = color == COLOR_Y ? cur_cu->log2_height : uvg_g_convert_to_log2[height];
  }
  return (unsigned)B200(quantize_residual)(&c, cur_cu->type, cur_cu->tr_idx, width, height, (int)color, use_trskip,
                                           in_stride, out_stride, ref_in, pred_in, rec_out, coeff_out, early_skip);
}

static uint32_t coeff_abs_sum_cuda(const coeff_t *coeffs, size_t length)
{ return uvg_b200_coeff_abs_sum(coeffs, length); }
static uint32_t fast_coeff_cost_cuda(const coeff_t *coeff, int32_t w, int32_t h, uint64_t weights)
{ return uvg_b200_fast_coeff_cost(coeff, w, h, weights); }

/* ---- intra (row 8(f)-2): strategies-intra.h:46-80; cu_loc_t is flattened to the w x h of the plane
 * block that is predicted, exactly what the generic leaves derive first (intra-generic.c:66-67,
 * :313-314, :421-422) ------------------------------------------------------------------- */
static void angular_pred_cuda(const cu_loc_t *const cu_loc, const int_fast8_t intra_mode, const int_fast8_t channel_type,
                              const uvg_pixel *const in_ref_above, const uvg_pixel *const in_ref_left,
                              uvg_pixel *const dst, const uint8_t multi_ref_idx, const uint8_t isp_mode, const int cu_dim)
{
  (void)cu_dim;
  const int w = channel_type == COLOR_Y ? cu_loc->width : cu_loc->chroma_width;
  const int h = channel_type == COLOR_Y ? cu_loc->height : cu_loc->chroma_height;
  B200(angular_pred)(w, h, intra_mode, channel_type, in_ref_above, in_ref_left, dst, multi_ref_idx, isp_mode);
}
static void intra_pred_planar_cuda(const cu_loc_t *const cu_loc, color_t color, const uvg_pixel *const ref_top,
                                   const uvg_pixel *const ref_left, uvg_pixel *const dst)
{
  const int w = color == COLOR_Y ? cu_loc->width : cu_loc->chroma_width;
  const int h = color == COLOR_Y ? cu_loc->height : cu_loc->chroma_height;
  B200(intra_pred_planar)(w, h, ref_top, ref_left, dst);
}
static void pdpc_planar_dc_cuda(const int mode, const cu_loc_t *const cu_loc, const color_t color,
                                const uvg_intra_ref *const used_ref, uvg_pixel *const dst)
{
  const int w = color == COLOR_Y ? cu_loc->width : cu_loc->chroma_width;
  const int h = color == COLOR_Y ? cu_loc->height : cu_loc->chroma_height;
  B200(pdpc_planar_dc)(mode, w, h, used_ref->top, used_ref->left, dst);
}

/* ---- ipol (row 8(f)-1): strategies-ipol.h:105-108, the MC sample-interpolation leaves ---------------- */
static void sample_quarterpel_luma_cuda(const encoder_control_t *const encoder, uvg_pixel *src, int16_t src_stride, int width,
                                        int height, uvg_pixel *dst, int16_t dst_stride, int8_t hor_flag, int8_t ver_flag,
                                        const mv_t mv[2])
{ (void)encoder; (void)hor_flag; (void)ver_flag; B200(sample_quarterpel_luma)(src, src_stride, width, height, dst, dst_stride, mv[0], mv[1]); }
static void sample_quarterpel_luma_hi_cuda(const encoder_control_t *const encoder, uvg_pixel *src, int16_t src_stride, int width,
                                           int height, int16_t *dst, int16_t dst_stride, int8_t hor_flag, int8_t ver_flag,
                                           const mv_t mv[2])
{ (void)encoder; (void)hor_flag; (void)ver_flag; B200(sample_quarterpel_luma_hi)(src, src_stride, width, height, dst, dst_stride, mv[0], mv[1]); }
static void sample_octpel_chroma_cuda(const encoder_control_t *const encoder, uvg_pixel *src, int16_t src_stride, int width,
                                      int height, uvg_pixel *dst, int16_t dst_stride, int8_t hor_flag, int8_t ver_flag,
                                      const mv_t mv[2])
{ (void)encoder; (void)hor_flag; (void)ver_flag; B200(sample_octpel_chroma)(src, src_stride, width, height, dst, dst_stride, mv[0], mv[1]); }
static void sample_octpel_chroma_hi_cuda(const encoder_control_t *const encoder, uvg_pixel *src, int16_t src_stride, int width,
                                         int height, int16_t *dst, int16_t dst_stride, int8_t hor_flag, int8_t ver_flag,
                                         const mv_t mv[2])
{ (void)encoder; (void)hor_flag; (void)ver_flag; B200(sample_octpel_chroma_hi)(src, src_stride, width, height, dst, dst_stride, mv[0], mv[1]); }

/* The four FME block filters (ipol_blocks_func, strategies-ipol.h:62-64): the outputs are plain sample
 * interpolations at fixed quarter-pel offsets, so the shared intermediate buffers are ignored. They are
 * always installed together. */
#define FME_STEP(name, step)                                                                                       \
  static void name(const encoder_control_t *encoder, uvg_pixel *src, int16_t src_stride, int width, int height,    \
                   uvg_pixel filtered[4][LCU_LUMA_SIZE], int16_t hor_intermediate[5][UVG_IPOL_MAX_IM_SIZE_LUMA_SIMD], \
                   int8_t fme_level, int16_t hor_first_cols[5][UVG_EXT_BLOCK_W_LUMA + 1], int8_t off_x, int8_t off_y) \
  {                                                                                                                \
    (void)encoder; (void)hor_intermediate; (void)fme_level; (void)hor_first_cols;                                  \
    B200(fme_filter)(step, src, src_stride, width, height, &filtered[0][0], LCU_WIDTH, LCU_LUMA_SIZE, off_x, off_y); \
  }
FME_STEP(filter_hpel_blocks_hor_ver_luma_cuda, 0)
FME_STEP(filter_hpel_blocks_diag_luma_cuda, 1)
FME_STEP(filter_qpel_blocks_hor_ver_luma_cuda, 2)
FME_STEP(filter_qpel_blocks_diag_luma_cuda, 3)

static void install_rough_hook(void);

/* ---- registration ---------------------------------------------------------------- */
static int device_ready(uint8_t bitdepth)
{
  if (bitdepth != UVG_BIT_DEPTH) {
    fprintf(stderr, "uvg266_b200: glue built for %d-bit pixels, encoder asks for %d\n", UVG_BIT_DEPTH, bitdepth);
    return 0;
  }
  const char *dev = getenv("UVG266_B200_DEVICE");
  if (uvg_b200_init(dev ? atoi(dev) : 0) != 0) {
    fprintf(stderr, "uvg266_b200: %s\n", uvg_b200_last_error());
    return 0;
  }
  uvg_b200_set_dropin_bitdepth(UVG_BIT_DEPTH);
  return 1;
}

/* Every wrapper is first assigned to a pointer of the reference's own typedef for that strategy type, so a
 * signature drift against the reference headers is a compile error here instead of corrupted arguments. */
#define REG(type, fntype, fn)                                                                    \
  do { fntype *const typed_ = fn;                                                                \
       ok &= uvg_strategyselector_register(opaque, type, STRATEGY_NAME, CUDA_PRIORITY, (void *)typed_); } while (0)

int uvg_b200_install_rough_hook_only(uint8_t bitdepth)
{
  if (!device_ready(bitdepth)) return 0;
  install_rough_hook();
  return 1;
}

int uvg_strategy_register_picture_cuda(void *opaque, uint8_t bitdepth)
{
  if (!device_ready(bitdepth)) return 0;
  int ok = 1;
  install_rough_hook();
  REG("reg_sad", reg_sad_func, reg_sad_cuda);
  REG("sad_4x4", cost_pixel_nxn_func, sad_4_cuda); REG("sad_8x8", cost_pixel_nxn_func, sad_8_cuda); REG("sad_16x16", cost_pixel_nxn_func, sad_16_cuda);
  REG("sad_32x32", cost_pixel_nxn_func, sad_32_cuda); REG("sad_64x64", cost_pixel_nxn_func, sad_64_cuda);
  REG("satd_4x4", cost_pixel_nxn_func, satd_4_cuda); REG("satd_8x8", cost_pixel_nxn_func, satd_8_cuda); REG("satd_16x16", cost_pixel_nxn_func, satd_16_cuda);
  REG("satd_32x32", cost_pixel_nxn_func, satd_32_cuda); REG("satd_64x64", cost_pixel_nxn_func, satd_64_cuda);
  REG("satd_any_size", cost_pixel_any_size_func, satd_any_size_cuda); REG("satd_any_size_vtm", cost_pixel_any_size_func, satd_any_size_vtm_cuda);
  REG("sad_4x4_dual", cost_pixel_nxn_multi_func, sad_4_dual_cuda); REG("sad_8x8_dual", cost_pixel_nxn_multi_func, sad_8_dual_cuda);
  REG("sad_16x16_dual", cost_pixel_nxn_multi_func, sad_16_dual_cuda); REG("sad_32x32_dual", cost_pixel_nxn_multi_func, sad_32_dual_cuda);
  REG("satd_4x4_dual", cost_pixel_nxn_multi_func, satd_4_dual_cuda); REG("satd_8x8_dual", cost_pixel_nxn_multi_func, satd_8_dual_cuda);
  REG("satd_16x16_dual", cost_pixel_nxn_multi_func, satd_16_dual_cuda); REG("satd_32x32_dual", cost_pixel_nxn_multi_func, satd_32_dual_cuda);
  /* unreachable from the encoder (search_intra.c:1000) but in the table: pred_buffer semantics kept
   * literally - candidate 1 starts 1024 pixels after candidate 0 and both read 4096 pixels */
  REG("sad_64x64_dual", cost_pixel_nxn_multi_func, sad_64_dual_cuda); REG("satd_64x64_dual", cost_pixel_nxn_multi_func, satd_64_dual_cuda);
  REG("satd_any_size_quad", cost_pixel_any_size_multi_func, satd_any_size_quad_cuda);
  REG("pixels_calc_ssd", pixels_calc_ssd_func, pixels_calc_ssd_cuda);
  REG("get_optimized_sad", get_optimized_sad_func, get_optimized_sad_cuda);
  REG("ver_sad", ver_sad_func, ver_sad_cuda); REG("hor_sad", hor_sad_func, hor_sad_cuda);
  REG("generate_residual", generate_residual_func, generate_residual_cuda);
  REG("bipred_average", inter_recon_bipred_func, bipred_average_cuda);
  REG("crc32c_4x4", crc32c_4x4_func, crc32c_4x4_cuda); REG("crc32c_8x8", crc32c_8x8_func, crc32c_8x8_cuda);
  REG("pixel_var", pixel_var_func, pixel_var_cuda);
  return ok;
}

int uvg_strategy_register_dct_cuda(void *opaque, uint8_t bitdepth)
{
  if (!device_ready(bitdepth)) return 0;
  int ok = 1;
  REG("dct_4x4", dct_func, dct_4x4_cuda); REG("dct_8x8", dct_func, dct_8x8_cuda); REG("dct_16x16", dct_func, dct_16x16_cuda);
  REG("dct_32x32", dct_func, dct_32x32_cuda);
  REG("idct_4x4", dct_func, idct_4x4_cuda); REG("idct_8x8", dct_func, idct_8x8_cuda); REG("idct_16x16", dct_func, idct_16x16_cuda);
  REG("idct_32x32", dct_func, idct_32x32_cuda);
  REG("fast_forward_dst_4x4", dct_func, fast_forward_dst_4x4_cuda); REG("fast_inverse_dst_4x4", dct_func, fast_inverse_dst_4x4_cuda);
  REG("mts_dct", mts_dct_func, mts_dct_cuda); REG("mts_idct", mts_idct_func, mts_idct_cuda);
  return ok;
}

/* ---- batched call site (row 8(f)-4): search_intra_rough through ONE launch ------------------------------------------
 * A reference built with glue/patches/search_intra_rough.patch exports the pointer uvg_b200_rough_hook; when it is
 * there the picture registrar installs this function: all 67 modes of a square 8 / 16 / 32 luma block predicted and
 * scored min(SATD, 2 SAD) (get_cost_dual, search_intra.c:133-160) by the fused rough-search kernel. Returns 0 for
 * blocks it does not cover (the encoder then walks its own mode-by-mode path). */
typedef int (*rough_hook_t)(const uvg_intra_references *refs, const uvg_pixel *orig_block, int width, int height,
                            unsigned *costs_out);
static int rough_hook_cuda(const uvg_intra_references *refs, const uvg_pixel *orig_block, int width, int height,
                           unsigned *costs_out)
{
#if UVG_BIT_DEPTH == 8
  static const int8_t all_modes[67] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25,
    26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56,
    57, 58, 59, 60, 61, 62, 63, 64, 65, 66 };
  if (width != height || (width != 8 && width != 16 && width != 32)) return 0;
  uvg_pixel rows[2 * (2 * 32 + 4)];
  const int gp = 2 * width + 4;
  memcpy(rows, refs->ref.top, (2 * width + 1) * sizeof(uvg_pixel));
  memcpy(rows + gp, refs->ref.left, (2 * width + 1) * sizeof(uvg_pixel));
  uvg_b200_blocks o;
  o.base = orig_block; o.stride = (uint32_t)width; o.blocks_per_row = 1; o.col_pitch = 0; o.row_pitch = (uint64_t)width * height;
  if (uvg_b200_intra_rough_host(8, width, &o, (size_t)width * height, rows, all_modes, 67, costs_out, 1) != 0) {
    fprintf(stderr, "uvg266_b200: %s\n", uvg_b200_last_error());
    abort();
  }
  return 1;
#else
  (void)refs; (void)orig_block; (void)width; (void)height; (void)costs_out;
  return 0;
#endif
}

static void install_rough_hook(void)
{
  rough_hook_t *slot = (rough_hook_t *)dlsym(RTLD_DEFAULT, "uvg_b200_rough_hook");
  const char *off = getenv("UVG266_B200_ROUGH_HOOK");
  if (slot && !(off && !strcmp(off, "0"))) *slot = rough_hook_cuda;
}

/* mip_predict (its trained matrices are file-static in mip_data_generic.h, not exported, so they cannot be handed to the
 * device at run time) and the unused intra_pred_filtered_dc stay with the reference's implementations */
int uvg_strategy_register_intra_cuda(void *opaque, uint8_t bitdepth)
{
  if (!device_ready(bitdepth)) return 0;
  int ok = 1;
  REG("angular_pred", angular_pred_func, angular_pred_cuda); REG("intra_pred_planar", intra_pred_planar_func, intra_pred_planar_cuda);
  REG("pdpc_planar_dc", pdpc_planar_dc_func, pdpc_planar_dc_cuda);
  return ok;
}

int uvg_strategy_register_ipol_cuda(void *opaque, uint8_t bitdepth)
{
  if (!device_ready(bitdepth)) return 0;
  int ok = 1;
  REG("sample_quarterpel_luma", uvg_sample_quarterpel_luma_func, sample_quarterpel_luma_cuda); REG("sample_quarterpel_luma_hi", uvg_sample_quarterpel_luma_hi_func, sample_quarterpel_luma_hi_cuda);
  REG("sample_octpel_chroma", uvg_sample_octpel_chroma_func, sample_octpel_chroma_cuda); REG("sample_octpel_chroma_hi", uvg_sample_octpel_chroma_hi_func, sample_octpel_chroma_hi_cuda);
  REG("filter_hpel_blocks_hor_ver_luma", ipol_blocks_func, filter_hpel_blocks_hor_ver_luma_cuda);
  REG("filter_hpel_blocks_diag_luma", ipol_blocks_func, filter_hpel_blocks_diag_luma_cuda);
  REG("filter_qpel_blocks_hor_ver_luma", ipol_blocks_func, filter_qpel_blocks_hor_ver_luma_cuda);
  REG("filter_qpel_blocks_diag_luma", ipol_blocks_func, filter_qpel_blocks_diag_luma_cuda);
  /* get_extended_block(_wraparound): pure memcpy / border replication on the host side, left alone */
  return ok;
}

/* quant_cbcr_residual (JCCR, strategies-quant.h:62-80): the plain chain is one fused kernel over both chroma
 * planes; with LFNST, RDOQ, dep-quant, LMCS or scaling lists in play the call is forwarded. */
static unsigned quant_cbcr_residual_cuda(encoder_state_t *const state, const cu_info_t *const cur_cu, const int width,
                                         const int height, const coeff_scan_order_t scan_order, const int in_stride,
                                         const int out_stride, const uvg_pixel *const u_ref_in,
                                         const uvg_pixel *const v_ref_in, const uvg_pixel *const u_pred_in,
                                         const uvg_pixel *const v_pred_in, uvg_pixel *u_rec_out, uvg_pixel *v_rec_out,
                                         coeff_t *coeff_out, bool early_skip, int lmcs_chroma_adj,
                                         enum uvg_tree_type tree_type)
{
  const encoder_control_t *const enc = state->encoder_control;
  const uint8_t lfnst_idx = tree_type == UVG_CHROMA_T ? cur_cu->cr_lfnst_idx : cur_cu->lfnst_idx;
  const bool rdoq = enc->cfg.rdoq_enable && (width > 4 || !enc->cfg.rdoq_skip);
  const bool lmcs = state->tile->frame->lmcs_aps->m_sliceReshapeInfo.enableChromaAdj;
  if (lfnst_idx || rdoq || enc->cfg.dep_quant || lmcs || scaling_lists_on(enc))
    return ref_quant_cbcr_residual(state, cur_cu, width, height, scan_order, in_stride, out_stride, u_ref_in, v_ref_in,
                                   u_pred_in, v_pred_in, u_rec_out, v_rec_out, coeff_out, early_skip, lmcs_chroma_adj,
                                   tree_type);
  uvg_b200_quant_cfg c;
  fill_cfg(state, cur_cu, &c);
  return (unsigned)B200(quant_cbcr_residual)(&c, cur_cu->type, cur_cu->tr_idx, width, height, cur_cu->joint_cb_cr,
                                             state->frame->jccr_sign ? 1 : 0, in_stride, out_stride, u_ref_in, v_ref_in,
                                             u_pred_in, v_pred_in, u_rec_out, v_rec_out, coeff_out, early_skip);
}

/* Highest-priority implementation of `type` registered so far that is not ours. */
static void *registered_before_us(const strategy_list_t *const list, const char *type)
{
  void *f = NULL;
  unsigned best = 0;
  for (unsigned i = 0; i < list->count; ++i) {
    const strategy_t *const st = &list->strategies[i];
    if (!strcmp(st->type, type) && strcmp(st->strategy_name, STRATEGY_NAME) && (!f || st->priority >= best)) {
      f = st->fptr;
      best = st->priority;
    }
  }
  return f;
}

int uvg_strategy_register_quant_cuda(void *opaque, uint8_t bitdepth)
{
  if (!device_ready(bitdepth)) return 0;
  int ok = 1;
  const strategy_list_t *const list = (const strategy_list_t *)opaque;
  ref_quant = (quant_func *)registered_before_us(list, "quant");
  ref_dequant = (dequant_func *)registered_before_us(list, "dequant");
  ref_quantize_residual = (quant_residual_func *)registered_before_us(list, "quantize_residual");
  ref_quant_cbcr_residual = (quant_cbcr_func *)registered_before_us(list, "quant_cbcr_residual");
  if (!ref_quant || !ref_dequant || !ref_quantize_residual || !ref_quant_cbcr_residual) {
    /* fail at registration, never mid-encode */
    fprintf(stderr, "uvg266_b200: the reference's quant strategies must be registered before the cuda ones\n");
    return 0;
  }
  /* The trained LFNST kernels are exported read-only data of the encoder itself (lfnst_tables.h:56, :211): hand the
   * library their addresses; without them (a build that hides the symbols) LFNST blocks are forwarded as before.
   * UVG266_B200_LFNST=host keeps LFNST on the host for A/B runs. */
  {
    const int8_t *t8 = (const int8_t *)dlsym(RTLD_DEFAULT, "uvg_lfnst_8x8");
    const int8_t *t4 = (const int8_t *)dlsym(RTLD_DEFAULT, "uvg_lfnst_4x4");
    const char *where = getenv("UVG266_B200_LFNST");
    lfnst_on_device = t8 && t4 && !(where && !strcmp(where, "host")) && uvg_b200_set_lfnst_tables(t8, t4) == 0;
  }
  REG("quant", quant_func, quant_cuda); REG("dequant", dequant_func, dequant_cuda);
  REG("quantize_residual", quant_residual_func, quantize_residual_cuda);
  REG("quant_cbcr_residual", quant_cbcr_func, quant_cbcr_residual_cuda);
  REG("coeff_abs_sum", coeff_abs_sum_func, coeff_abs_sum_cuda); REG("fast_coeff_cost", fast_coeff_cost_func, fast_coeff_cost_cuda);
  return ok;
}
