/*
 * uvg_oracle.h - TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of uvg266's block-distortion + transform + quant leaf
 * functions, i.e. of the *generic* strategies in
 *   /root/reference/src/strategies/generic/picture-generic.c
 *   /root/reference/src/strategies/generic/dct-generic.c
 *   /root/reference/src/strategies/generic/quant-generic.c
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it; the
 * product library (uvg266_b200/lib/libuvg266_b200.so) never links or calls it.
 *
 * Parity status: PINNED. tests/test_oracle_vs_ref.py checks every function below against
 * the reference's own golden vectors (tests/satd_tests.c:122-159, tests/sad_tests.c:144-282,
 * tests/intra_sad_tests.c, tests/coeff_sum_tests.c) and, on seeded random inputs, against
 * the unmodified reference compiled into oracle/_ref/ (strategy "generic").
 *
 * Pixel type is compile-time like the reference's UVG_BIT_DEPTH (uvg266.h:89-97):
 * build with -DORC_BIT_DEPTH=8 (uint8 pixels) or 10 (uint16 pixels).
 */
#ifndef UVG_ORACLE_H_
#define UVG_ORACLE_H_
#include <stdint.h>
#include <stddef.h>

#ifndef ORC_BIT_DEPTH
#define ORC_BIT_DEPTH 8
#endif
#if ORC_BIT_DEPTH == 8
typedef uint8_t orc_pixel;
#else
typedef uint16_t orc_pixel;
#endif

#ifdef __cplusplus
extern "C" {
#endif

int orc_bit_depth(void);

/* ---- strategies-picture ---------------------------------------------------------- */
unsigned orc_reg_sad(const orc_pixel *d1, const orc_pixel *d2, int w, int h,
                     unsigned stride1, unsigned stride2);
unsigned orc_sad_nxn(int n, const orc_pixel *b1, const orc_pixel *b2);
void orc_sad_nxn_dual(int n, const orc_pixel *preds /* 2 slots of 1024 px */,
                      const orc_pixel *orig, unsigned *costs_out);
unsigned orc_satd_4x4(const orc_pixel *b1, const orc_pixel *b2);
unsigned orc_satd_nxn(int n, const orc_pixel *b1, const orc_pixel *b2);
void orc_satd_nxn_dual(int n, const orc_pixel *preds, const orc_pixel *orig, unsigned *costs_out);
unsigned orc_satd_any_size(int w, int h, const orc_pixel *b1, int stride1,
                           const orc_pixel *b2, int stride2);
void orc_satd_any_size_quad(int w, int h, const orc_pixel *const preds[4], int stride,
                            const orc_pixel *orig, int orig_stride, unsigned *costs_out);
unsigned orc_satd_any_size_vtm(int w, int h, const orc_pixel *org, int org_stride,
                               const orc_pixel *cur, int cur_stride);
unsigned orc_pixels_calc_ssd(const orc_pixel *ref, const orc_pixel *rec, int ref_stride,
                             int rec_stride, int w, int h);
uint32_t orc_ver_sad(const orc_pixel *pic, const orc_pixel *ref_row, int w, int h,
                     uint32_t pic_stride);
uint32_t orc_hor_sad(const orc_pixel *pic, const orc_pixel *ref, int w, int h,
                     uint32_t pic_stride, uint32_t ref_stride, uint32_t left, uint32_t right);
void orc_generate_residual(const orc_pixel *ref, const orc_pixel *pred, int16_t *residual,
                           int w, int h, int ref_stride, int pred_stride);
/* uvg_image_calc_sad semantics (image.c:438): SAD against the edge-replicated reference. */
unsigned orc_image_calc_sad(const orc_pixel *pic, const orc_pixel *ref, int pic_w, int pic_h,
                            int stride, int pic_x, int pic_y, int ref_x, int ref_y,
                            int bw, int bh);

/* ---- strategies-dct -------------------------------------------------------------- */
enum { ORC_DCT2 = 0, ORC_DCT8 = 1, ORC_DST7 = 2 };            /* tr_type_t, uvg266.h:234 */
enum { ORC_MTS_OFF = 0, ORC_MTS_INTRA = 1, ORC_MTS_INTER = 2, ORC_MTS_BOTH = 3,
       ORC_MTS_IMPLICIT = 4 };                                 /* uvg_mts, uvg266.h */
enum { ORC_CU_INTRA = 1, ORC_CU_INTER = 2 };                   /* cu_type_t, cu.h:50 */

/* n*n row-major matrix T[k*n+i] of the given type (n in 2,4,8,16,32; DCT2 only for 2). */
int orc_tr_matrix(int type, int n, int16_t *out);
void orc_dct_nxn(int n, int bitdepth, const int16_t *in, int16_t *out);
void orc_idct_nxn(int n, int bitdepth, const int16_t *in, int16_t *out);
void orc_fast_forward_dst_4x4(int bitdepth, const int16_t *in, int16_t *out);
void orc_fast_inverse_dst_4x4(int bitdepth, const int16_t *in, int16_t *out);
void orc_get_tr_type(int w, int h, int color, int cu_type, int tr_idx, int lfnst_idx,
                     int cr_lfnst_idx, int isp_mode, int mts_type, int *hor, int *ver);
void orc_mts_dct(int bitdepth, int color, int cu_type, int tr_idx, int lfnst_idx,
                 int cr_lfnst_idx, int isp_mode, int w, int h, const int16_t *in,
                 int16_t *out, int mts_type);
void orc_mts_idct(int bitdepth, int color, int cu_type, int tr_idx, int lfnst_idx,
                  int cr_lfnst_idx, int isp_mode, int w, int h, const int16_t *in,
                  int16_t *out, int mts_type);

/* ---- strategies-quant ------------------------------------------------------------ */
typedef struct {
  int32_t qp;              /* state->qp                                  */
  int32_t bitdepth;        /* encoder_control->bitdepth                  */
  int32_t slice_intra;     /* state->frame->slicetype == UVG_SLICE_I     */
  int32_t signhide;        /* cfg.signhide_enable                        */
  int32_t mts;             /* cfg.mts (selects mts_dct in transform2d)   */
  const int8_t *chroma_qp_map; /* encoder_control->qp_map[0], or NULL    */
  int32_t isp_mode;        /* cur_cu->intra.isp_mode (uvg_get_tr_type, dct-generic.c:2522) */
} orc_quant_params;

/* scan[(w*h)] of uvg_get_scan_order_table(SCAN_GROUP_4X4, SCAN_DIAG, log2w, log2h, 0) */
void orc_scan_table(int log2_w, int log2_h, uint32_t *out);
int32_t orc_get_scaled_qp(int color, int qp, int qp_offset, const int8_t *chroma_map);
void orc_quant(const orc_quant_params *p, const int16_t *coef, int16_t *q_coef, int w, int h,
               int color, int block_type, int transform_skip, int lfnst_idx);
void orc_dequant(const orc_quant_params *p, const int16_t *q_coef, int16_t *coef, int w, int h,
                 int color, int block_type, int transform_skip);
/* residual -> T -> Q -> has_coeffs -> IQ -> IT -> rec (flat lists, no lfnst/rdoq/dep-quant,
 * no LMCS): uvg_quantize_residual_generic, quant-generic.c:460-612 */
int orc_quantize_residual(const orc_quant_params *p, int cu_type, int tr_idx, int w, int h,
                          int color, int use_trskip, int in_stride, int out_stride,
                          const orc_pixel *ref_in, const orc_pixel *pred_in,
                          orc_pixel *rec_out, int16_t *coeff_out, int early_skip);
/* LFNST (transform.c:878-1200). The trained kernels come from the compiled reference at run time. */
typedef struct {
  int lfnst_idx, cr_lfnst_idx;      /* cu_info_t.lfnst_idx / cr_lfnst_idx */
  int intra_mode, mode_chroma;      /* cu_info_t.intra.mode / mode_chroma */
  int cu_log2_w, cu_log2_h;         /* cu_info_t.log2_width / log2_height */
  int tree_type;                    /* 0 both, 1 luma, 2 chroma (cu.h:91-93) */
  int luma_mode;                    /* state->collocated_luma_mode (CCLM substitution) */
} orc_lfnst_info;
void orc_set_lfnst_tables(const int8_t *t8x8, const int8_t *t4x4);
void orc_fwd_lfnst(int w, int h, int lfnst_idx, int intra_mode, int wa_log2_w, int wa_log2_h, int16_t *coeffs);
void orc_inv_lfnst(int w, int h, int lfnst_idx, int intra_mode, int wa_log2_w, int wa_log2_h, int16_t *coeffs);
int orc_quantize_residual_lfnst(const orc_quant_params *p, const orc_lfnst_info *li, int cu_type, int tr_idx,
                                int w, int h, int color, int in_stride, int out_stride,
                                const orc_pixel *ref_in, const orc_pixel *pred_in, orc_pixel *rec_out,
                                int16_t *coeff_out, int early_skip);
int orc_quant_cbcr_residual(const orc_quant_params *p, int cu_type, int tr_idx, int w, int h, int joint_cb_cr,
                            int jccr_sign, int in_stride, int out_stride, const orc_pixel *u_ref,
                            const orc_pixel *v_ref, const orc_pixel *u_pred, const orc_pixel *v_pred,
                            orc_pixel *u_rec, orc_pixel *v_rec, int16_t *coeff_out, int early_skip);
uint32_t orc_coeff_abs_sum(const int16_t *coeffs, size_t length);
uint32_t orc_fast_coeff_cost(const int16_t *coeff, int w, int h, uint64_t weights);

/* ---- strategies-intra (next row 8(f)-2), luma, mrl = 0, no ISP / MIP ----------------- */
void orc_angular_pred(int w, int h, int pred_mode, const orc_pixel *ref_above, const orc_pixel *ref_left,
                      orc_pixel *dst);
void orc_angular_pred_leaf(int w, int h, int intra_mode, int channel_type, const orc_pixel *ref_above,
                           const orc_pixel *ref_left, orc_pixel *dst, int multi_ref_idx, int isp_mode);
void orc_intra_pred_planar(int w, int h, const orc_pixel *ref_top, const orc_pixel *ref_left, orc_pixel *dst);
void orc_intra_pred_dc(int w, int h, const orc_pixel *ref_top, const orc_pixel *ref_left, orc_pixel *dst);
void orc_pdpc_planar_dc(int w, int h, const orc_pixel *ref_top, const orc_pixel *ref_left, orc_pixel *dst);
void orc_intra_filter_reference(int w, int h, const orc_pixel *top, const orc_pixel *left, orc_pixel *ftop,
                                orc_pixel *fleft);
void orc_intra_predict(int w, int h, int mode, const orc_pixel *top, const orc_pixel *left, orc_pixel *dst);

/* ---- strategies-ipol (next row 8(f)-1): MC sample interpolation ------------------------------ */
void orc_sample_quarterpel_luma(int w, int h, const orc_pixel *src, int src_stride, orc_pixel *dst, int dst_stride,
                                int mvx, int mvy);
void orc_sample_quarterpel_luma_hi(int w, int h, const orc_pixel *src, int src_stride, int16_t *dst, int dst_stride,
                                   int mvx, int mvy);
void orc_sample_octpel_chroma(int w, int h, const orc_pixel *src, int src_stride, orc_pixel *dst, int dst_stride,
                              int mvx, int mvy);
void orc_sample_octpel_chroma_hi(int w, int h, const orc_pixel *src, int src_stride, int16_t *dst, int dst_stride,
                                 int mvx, int mvy);

void orc_fme_filter(int step, const orc_pixel *src, int src_stride, int w, int h, int off_x, int off_y, orc_pixel *out);

double orc_pixel_var(const orc_pixel *arr, uint32_t len);
/* ---- picture table leftovers: IBC hash and bi-prediction average ---------------------------------- */
uint32_t orc_crc32c_nxn(int n, const orc_pixel *buf, uint32_t stride);
void orc_bipred_average(orc_pixel *dst, int dst_stride, const void *src0, int src0_is_im, const void *src1,
                        int src1_is_im, int w, int h);

#ifdef __cplusplus
}
#endif
#endif
