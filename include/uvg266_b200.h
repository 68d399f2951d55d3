/*
 * uvg266_b200.h - C ABI of libuvg266_b200.so: B200 (sm_100a) implementation of uvg266's
 * block-distortion + transform + quant hot path.
 *
 * Three layers, all plain C (pointers + sizes, no CUDA/torch types in any signature):
 *
 *  (1) DROP-IN leaf functions with exactly the reference's strategy signatures
 *      (host pointers, synchronous, re-entrant from any number of encoder threads). These
 *      are what uvg266's strategyselector table binds; see INTEGRATION.md for the
 *      registration glue (uvg266_b200/glue/strategies_cuda.c).
 *        strategies-picture.h:115-157   reg_sad, sad/satd NxN (+dual), satd_any_size(+quad,
 *                                       +vtm), pixels_calc_ssd, ver_sad, hor_sad,
 *                                       generate_residual, get_optimized_sad
 *        strategies-dct.h:44,63-83      dct/idct NxN, fast dst 4x4, mts_dct, mts_idct
 *        strategies-quant.h:48-91       quant, dequant, quantize_residual, coeff_abs_sum,
 *                                       fast_coeff_cost
 *      Reference-private structs (encoder_state_t, cu_info_t) do not cross this ABI: the
 *      glue flattens them into uvg_b200_quant_cfg / plain ints.
 *
 *  (2) BATCHED DEVICE entry points (`*_dev`): device pointers, asynchronous on a caller
 *      stream. One launch evaluates thousands of blocks / TUs; this is the path the
 *      roofline numbers are measured on.
 *
 *  (3) BATCHED HOST entry points (`*_host`): same work from host buffers; H2D, kernels and
 *      D2H are issued inside the call (the end-to-end number of bench.py).
 *
 * Pixel type follows the reference's compile-time UVG_BIT_DEPTH (uvg266.h:89-97): every
 * function taking pixels has a `bitdepth` argument (8 -> uint8_t pixels, 9..16 ->
 * uint16_t pixels); the drop-in layer is instantiated per bit depth by the glue.
 *
 * All results are bit-exact with the reference's generic-C strategies.
 * Errors: functions returning int return 0 on success, non-zero on failure
 * (uvg_b200_last_error() has the text). Leaf drop-ins have no error channel
 * (strategies-*.h): a device failure after init aborts the process with a message -
 * there is NO CPU fallback anywhere in this library.
 */
#ifndef UVG266_B200_H_
#define UVG266_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UVG_B200_API __attribute__((visibility("default")))

/* ---------------------------------------------------------------------------------- */
/* Library / device                                                                    */
/* ---------------------------------------------------------------------------------- */
UVG_B200_API int uvg_b200_init(int device);           /* idempotent; selects the device  */
UVG_B200_API int uvg_b200_device_count(void);         /* 0 when no CUDA device/driver     */
UVG_B200_API const char *uvg_b200_last_error(void);
UVG_B200_API const char *uvg_b200_version(void);
UVG_B200_API int uvg_b200_sm_count(void);
/* Kernel launches issued by this library since load (all threads). bench.py reads it to
 * report gpu_launches. */
UVG_B200_API uint64_t uvg_b200_launch_count(void);
/* Launches of the tcgen05 + tensor-map-TMA rough-cost kernel (rough_tma.cu) since load. */
UVG_B200_API uint64_t uvg_b200_rough_tma_launches(void);

/* ---------------------------------------------------------------------------------- */
/* Block addressing for the batched entry points                                       */
/* ---------------------------------------------------------------------------------- */
/* Block i of an operand starts at pixel offset
 *     (i / blocks_per_row) * row_pitch + (i % blocks_per_row) * col_pitch
 * from `base`, and its rows are `stride` pixels apart. Two common instances:
 *   dense array of w*h blocks : stride = w, blocks_per_row = 1, row_pitch = w*h
 *   tiling of a picture plane : stride = plane stride, blocks_per_row = plane_w / w,
 *                               col_pitch = w, row_pitch = h * plane stride
 * `base` is a device pointer for *_dev functions and a host pointer for *_host ones. */
typedef struct uvg_b200_blocks {
  const void *base;
  uint32_t stride;
  uint32_t blocks_per_row;
  uint64_t col_pitch;
  uint64_t row_pitch;
} uvg_b200_blocks;

/* Distortion metric selector for the batched picture kernels. */
enum uvg_b200_metric {
  UVG_B200_SAD = 0,       /* reg_sad: sum|a-b|, no shift                    (picture-generic.c:99)   */
  UVG_B200_SAD_NXN = 1,   /* sad_NxN: sum|a-b| >> (bitdepth-8)              (:1054)                  */
  UVG_B200_SATD = 2,      /* satd_any_size / satd_NxN (N>=8): 8x8 + 4x4 Hadamards, >> (bitdepth-8)   */
  UVG_B200_SATD_4X4 = 3,  /* satd_4x4: one 4x4 Hadamard, no shift           (:205)                   */
  UVG_B200_SATD_VTM = 4,  /* satd_any_size_vtm (xGetHADs)                   (:949)                   */
  UVG_B200_SSD = 5,       /* pixels_calc_ssd: sum (a-b)^2 >> 2*(bitdepth-8) (:1115)                  */
  UVG_B200_SATD_QUAD = 6  /* per-candidate satd_any_size_quad semantics     (:412)                   */
};

/* ---------------------------------------------------------------------------------- */
/* (2) Batched device entry points                                                     */
/* ---------------------------------------------------------------------------------- */
/* out[i*k + c] = metric(block i of `a`, block (i*k + c) of `b`), i < n_blocks, c < k.
 * k = 1: independent pairs. k = 2 / 4: the reference's *_dual / *_quad shape. k = number
 * of intra modes: one original scored against all candidate predictions in one launch. */
UVG_B200_API int uvg_b200_distortion_dev(int metric, int bitdepth, int w, int h,
                                         const uvg_b200_blocks *a, const uvg_b200_blocks *b,
                                         uint32_t *out, size_t n_blocks, int k, void *stream);

/* Intra rough-search cost, search_intra.c:133-192 (get_cost_dual):
 * out[i*k + c] = min(satd, 2*sad) of square n*n blocks (satd_NxN_dual + sad_NxN_dual). */
UVG_B200_API int uvg_b200_rough_cost_dev(int bitdepth, int n, const uvg_b200_blocks *orig,
                                         const uvg_b200_blocks *preds, uint32_t *out,
                                         size_t n_blocks, int k, void *stream);

/* Motion-search SAD (search_inter.c:204-247 check_mv_cost -> image.c:438
 * uvg_image_calc_sad): for block j of the current picture (top-left cur_x[j], cur_y[j],
 * size w*h) and every integer displacement (dx,dy) in [-range,range]^2 around
 * (mv_x[j], mv_y[j]), out[j*(2r+1)^2 + (dy+r)*(2r+1) + (dx+r)] = SAD >> (bitdepth-8)
 * against the reference picture with frame-edge replication (ver/hor/cor_sad). */
UVG_B200_API int uvg_b200_sad_search_dev(int bitdepth, int w, int h,
                                         const void *cur, const void *ref, int pic_w, int pic_h,
                                         int stride, const int32_t *cur_x, const int32_t *cur_y,
                                         const int32_t *mv_x, const int32_t *mv_y, int range,
                                         uint32_t *out, size_t n_blocks, void *stream);

/* residual[i][y*w+x] = a - b (int16, dense), generate_residual (picture-generic.c:1360). */
UVG_B200_API int uvg_b200_residual_dev(int bitdepth, int w, int h, const uvg_b200_blocks *ref,
                                       const uvg_b200_blocks *pred, int16_t *residual,
                                       size_t n_blocks, void *stream);

/* Transform types, tr_type_t (uvg266.h:234). */
enum { UVG_B200_DCT2 = 0, UVG_B200_DCT8 = 1, UVG_B200_DST7 = 2 };

/* 2-D separable transform of n_tus dense w*h int16 blocks (mts_dct_generic /
 * mts_idct_generic, dct-generic.c:2560-2678, and dct_NxN/idct_NxN :720-740 which are the
 * DCT2/DCT2 instance). skip_w/skip_h: zero-out widths as derived by the reference
 * (:2582-2600). inverse = 0 forward, 1 inverse. */
UVG_B200_API int uvg_b200_transform_dev(int inverse, int bitdepth, int w, int h, int type_hor,
                                        int type_ver, int skip_w, int skip_h, const int16_t *in,
                                        int16_t *out, size_t n_tus, void *stream);

/* Quantiser configuration: the fields uvg_quant_generic / uvg_dequant_generic /
 * uvg_quantize_residual_generic read through encoder_state_t (last row). */
typedef struct uvg_b200_quant_cfg {
  int32_t qp;            /* state->qp                                                  */
  int32_t bitdepth;      /* encoder_control->bitdepth                                  */
  int32_t slice_intra;   /* state->frame->slicetype == UVG_SLICE_I                     */
  int32_t signhide;      /* cfg.signhide_enable                                        */
  int32_t mts;           /* cfg.mts (enum uvg_mts): selects mts_dct in uvg_transform2d */
  int32_t isp_mode;      /* cur_cu->intra.isp_mode of the CU the TU belongs to (0 for inter CUs): with
                          * explicit MTS an ISP luma TU uses DST-7 (uvg_get_tr_type, dct-generic.c:2522) */
  int8_t chroma_qp_map[128]; /* encoder_control->qp_map[0][qp] for qp 0..127 (entry =   */
  int32_t has_chroma_qp_map; /*   mapped qp), valid when has_chroma_qp_map != 0        */
  /* LFNST (cfg.lfnst, intra CUs; uvg_quantize_residual_generic quant-generic.c:503-509, :561-563; uvg_fwd_lfnst /
   * uvg_inv_lfnst transform.c:965-1200). All zero = LFNST off. Read by quantize_residual only.               */
  int32_t cu_lfnst_idx;      /* cur_cu->lfnst_idx    } the transform's zero-out depends on the CU fields      */
  int32_t cu_cr_lfnst_idx;   /* cur_cu->cr_lfnst_idx }   (mts_dct_generic, dct-generic.c:2589-2600)            */
  int32_t lfnst_index;       /* the index of this colour / tree (:503): what quant is told (first 8 / 16 scan   */
                             /*   positions only)                                                               */
  int32_t lfnst_apply;       /* uvg_fwd_lfnst's own gate (transform.c:990): lfnst_index && !mts_skip &&         */
                             /*   (luma || separate tree)                                                       */
  int32_t lfnst_mode;        /* the intra mode that picks the kernel: luma / chroma mode after the CCLM         */
                             /*   (-> collocated luma mode) and MIP (-> planar) substitutions (:1001-1006)      */
  int32_t lfnst_wa_log2_w;   /* block dims of the wide-angle correction (:1009-1014): cur_cu->log2_width/height  */
  int32_t lfnst_wa_log2_h;   /*   for luma, the TU's for chroma                                                  */
} uvg_b200_quant_cfg;

/* The two trained LFNST kernels, uvg_lfnst_8x8[4][2][16][48] and uvg_lfnst_4x4[4][2][16][16] (lfnst_tables.h:56,
 * :211): HOST pointers to the encoder's own exported read-only tables, copied to the device once. This library does
 * not contain them. Until it is called, quantize_residual with lfnst_apply != 0 fails. */
UVG_B200_API int uvg_b200_set_lfnst_tables(const int8_t *t8x8, const int8_t *t4x4);

/* q_coef = quant(coef) for n_tus dense w*h blocks (uvg_quant_generic, flat lists). */
UVG_B200_API int uvg_b200_quant_dev(const uvg_b200_quant_cfg *cfg, const int16_t *coef,
                                    int16_t *q_coef, int w, int h, int color,
                                    int transform_skip, int lfnst_idx, size_t n_tus,
                                    void *stream);
UVG_B200_API int uvg_b200_dequant_dev(const uvg_b200_quant_cfg *cfg, const int16_t *q_coef,
                                      int16_t *coef, int w, int h, int color,
                                      int transform_skip, size_t n_tus, void *stream);

/* The fused op of the path (uvg_quantize_residual_generic, quant-generic.c:460-612), one
 * launch for n_tus TUs of one shape:
 *   residual = ref - pred -> forward transform -> quant -> has_coeffs
 *   -> dequant -> inverse transform -> rec = clip(residual' + pred)   (rec = pred if none)
 * plus what the caller computes on the result right after (search.c:438, rdo.c:440):
 *   ssd[i]  = pixels_calc_ssd(ref, rec)            (may be NULL)
 *   cost[i] = fast_coeff_cost(coeff, weights)      (may be NULL)
 * coeff_out: dense int16 [n_tus][w*h]; has_coeffs: uint8 [n_tus]. rec may alias pred. */
UVG_B200_API int uvg_b200_quantize_residual_dev(const uvg_b200_quant_cfg *cfg, int cu_type,
                                                int tr_idx, int w, int h, int color,
                                                int use_trskip, int early_skip,
                                                const uvg_b200_blocks *ref,
                                                const uvg_b200_blocks *pred,
                                                const uvg_b200_blocks *rec,
                                                int16_t *coeff_out, uint8_t *has_coeffs,
                                                uint32_t *ssd, uint32_t *cost,
                                                uint64_t cost_weights, size_t n_tus,
                                                void *stream);

/* Joint Cb-Cr residual coding (quant_cbcr_residual, strategies-quant.h:62-80;
 * uvg_quant_cbcr_residual_generic, quant-generic.c:241-442) for the plain chain (no LFNST / RDOQ /
 * dep-quant), n_tus TUs per launch: the two chroma residuals are mixed into one
 * ((4cb +- 2cr)/5, (cb +- cr)/2, (4cr +- 2cb)/5 by joint_cb_cr = 2, 3, 1 and jccr_sign; C division),
 * transformed / quantised / dequantised / inverse transformed as plane V (joint_cb_cr == 1) or U,
 * split back (:388-429) and both planes reconstructed; without coefficients (or early_skip) both
 * reconstructions are the predictions. has_coeffs[i] = 0 / 1 (the strategy returns joint_cb_cr or 0). */
UVG_B200_API int uvg_b200_quant_cbcr_residual_dev(const uvg_b200_quant_cfg *cfg, int cu_type,
                                                  int tr_idx, int w, int h, int joint_cb_cr,
                                                  int jccr_sign, int early_skip,
                                                  const uvg_b200_blocks *u_ref,
                                                  const uvg_b200_blocks *v_ref,
                                                  const uvg_b200_blocks *u_pred,
                                                  const uvg_b200_blocks *v_pred,
                                                  const uvg_b200_blocks *u_rec,
                                                  const uvg_b200_blocks *v_rec,
                                                  int16_t *coeff_out, uint8_t *has_coeffs,
                                                  size_t n_tus, void *stream);

/* coeff_abs_sum / fast_coeff_cost over n_tus dense blocks of `len` coefficients. */
UVG_B200_API int uvg_b200_coeff_abs_sum_dev(const int16_t *coeffs, size_t len, uint32_t *out,
                                            size_t n_tus, void *stream);
UVG_B200_API int uvg_b200_fast_coeff_cost_dev(const int16_t *coeffs, int w, int h,
                                              uint64_t weights, uint32_t *out, size_t n_tus,
                                              void *stream);

/* ---- intra prediction (next row 8(f)-2): luma, multi_ref_idx 0, no ISP / MIP ------------------
 * Replaces, for square n x n blocks (n = 4, 8, 16, 32), uvg_intra_predict -> intra_predict_regular
 * (intra.c:660-753) and the strategies it dispatches to: uvg_angular_pred, uvg_intra_pred_planar,
 * uvg_pdpc_planar_dc (strategies-intra.h:47-93, intra-generic.c:55-437) plus the reference
 * smoothing (intra.c:190-225) and DC (intra.c:236-273).
 * refs: per block, top[2n+1] then left[2n+1] (index 0 = corner sample, the layout of
 *       uvg_intra_ref, intra.h:47-58), each row with a pitch of 2n+4 pixels; the samples past 2n are
 *       ignored (the kernel continues the row with the last sample like intra.c:879-886).
 * modes: k luma modes (0 planar, 1 DC, 2..66 angular), a HOST array, shared by all blocks.
 * preds: dense [n_blocks][k][n*n].            costs: [n_blocks][k] = min(SATD, 2*SAD) against orig
 *       (search_intra_rough's cost, search_intra.c:1055-1062 + get_cost :149-163).             */
UVG_B200_API int uvg_b200_intra_predict_dev(int bitdepth, int n, const void *refs,
                                            const int8_t *modes, int k, void *preds,
                                            size_t n_blocks, void *stream);
UVG_B200_API int uvg_b200_intra_rough_dev(int bitdepth, int n, const uvg_b200_blocks *orig,
                                          const void *refs, const int8_t *modes, int k,
                                          uint32_t *costs, size_t n_blocks, void *stream);

/* ---- motion-compensation sample interpolation (next row 8(f)-1) ------------------------------
 * Replaces uvg_sample_quarterpel_luma(_hi) / uvg_sample_octpel_chroma(_hi) (strategies-ipol.h:105-108,
 * generic/ipol-generic.c:134-210, :681-759) for a batch of blocks read straight from the reference
 * picture plane; samples outside the plane are the replicated border (what uvg_get_extended_block
 * builds, ipol-generic.c:761-811).
 * chroma 0: 8 taps, MV in 1/16 luma samples; chroma 1: 4 taps, MV in 1/32 chroma samples.
 * hi 0: dst = pixels (uni-prediction rounding); hi 1: dst = int16 14-bit intermediates (bi-pred).
 * bx, by, mvx, mvy: DEVICE arrays [n_blocks]: block origin in the plane and its motion vector.
 * dst: dense [n_blocks][h*w]. */
UVG_B200_API int uvg_b200_ipol_dev(int bitdepth, int chroma, int hi, const void *plane, int stride,
                                   int plane_w, int plane_h, int w, int h, const int32_t *bx,
                                   const int32_t *by, const int32_t *mvx, const int32_t *mvy,
                                   void *dst, size_t n_blocks, void *stream);
/* Fused fractional motion search (search_inter.c:1085-1250 search_frac: interpolate the sub-pel
 * neighbours of the best integer MV, score them with satd_any_size): for every block, k candidates
 * mv + offsets[c] (HOST array [k][2], 1/16 luma samples, each in -16..15), 8-bit luma, w and h
 * multiples of 8: out[b*k + c] = satd_any_size(original block, interpolated block). */
UVG_B200_API int uvg_b200_frac_search_dev(int bitdepth, int w, int h, const void *cur,
                                          const void *ref, int stride, int plane_w, int plane_h,
                                          const int32_t *bx, const int32_t *by, const int32_t *mvx,
                                          const int32_t *mvy, const int8_t *offsets, int k,
                                          uint32_t *out, size_t n_blocks, void *stream);

/* ---- the non-distortion leftovers of the picture table -------------------------------------------
 * crc32c_4x4 / crc32c_8x8 (strategies-picture.h:157-158, picture-generic.c:1371-1443): the IBC block
 * hash of n x n blocks addressed by a blocks descriptor (e.g. every block position of a plane).
 * bipred_average (strategies-picture.h:136-148, picture-generic.c:1133-1247), one plane at a time:
 * src0 / src1 dense [n_blocks][w*h], each either pixels (is_im 0) or int16 14-bit intermediates
 * (is_im 1, the output of uvg_b200_ipol_dev with hi = 1); dst through a blocks descriptor. */
UVG_B200_API int uvg_b200_crc32c_dev(int bitdepth, int n, const uvg_b200_blocks *blocks,
                                     uint32_t *out, size_t n_blocks, void *stream);
UVG_B200_API int uvg_b200_bipred_average_dev(int bitdepth, int w, int h,
                                             const uvg_b200_blocks *dst, const void *src0,
                                             int src0_is_im, const void *src1, int src1_is_im,
                                             size_t n_blocks, void *stream);

/* pixel_var (strategies-picture.h:150, picture-generic.c:1334-1357): variance of `len` pixels in double, squared
 * deviations accumulated in index order like the reference (bit-exact doubles). Array a = arrays + a * pitch pixels;
 * out: DEVICE array of n_arrays doubles. One CTA per array (the encoder calls it per picture and per 64x64 CTU). */
UVG_B200_API int uvg_b200_pixel_var_dev(int bitdepth, const void *arrays, size_t pitch, uint32_t len,
                                        double *out, size_t n_arrays, void *stream);

/* ---------------------------------------------------------------------------------- */
/* (3) Batched host entry points (H2D + kernels + D2H inside the call, synchronous)    */
/* ---------------------------------------------------------------------------------- */
UVG_B200_API int uvg_b200_distortion_host(int metric, int bitdepth, int w, int h,
                                          const uvg_b200_blocks *a, size_t a_bytes,
                                          const uvg_b200_blocks *b, size_t b_bytes,
                                          uint32_t *out, size_t n_blocks, int k);
UVG_B200_API int uvg_b200_rough_cost_host(int bitdepth, int n, const uvg_b200_blocks *orig,
                                          size_t orig_bytes, const uvg_b200_blocks *preds,
                                          size_t preds_bytes, uint32_t *out, size_t n_blocks,
                                          int k);
UVG_B200_API int uvg_b200_intra_predict_host(int bitdepth, int n, const void *refs,
                                             const int8_t *modes, int k, void *preds,
                                             size_t n_blocks);
UVG_B200_API int uvg_b200_intra_rough_host(int bitdepth, int n, const uvg_b200_blocks *orig,
                                           size_t orig_bytes, const void *refs,
                                           const int8_t *modes, int k, uint32_t *costs,
                                           size_t n_blocks);
UVG_B200_API int uvg_b200_transform_host(int inverse, int bitdepth, int w, int h, int type_hor,
                                         int type_ver, int skip_w, int skip_h,
                                         const int16_t *in, int16_t *out, size_t n_tus);
UVG_B200_API int uvg_b200_quantize_residual_host(const uvg_b200_quant_cfg *cfg, int cu_type,
                                                 int tr_idx, int w, int h, int color,
                                                 int use_trskip, int early_skip,
                                                 const uvg_b200_blocks *ref, size_t ref_bytes,
                                                 const uvg_b200_blocks *pred, size_t pred_bytes,
                                                 const uvg_b200_blocks *rec, size_t rec_bytes,
                                                 int16_t *coeff_out, uint8_t *has_coeffs,
                                                 uint32_t *ssd, uint32_t *cost,
                                                 uint64_t cost_weights, size_t n_tus);

/* ---------------------------------------------------------------------------------- */
/* (4) Frame pipeline: the batched call site, device-resident        */
/* ---------------------------------------------------------------------------------- */
/* What the encoder does per CTU of an all-intra picture, for a whole picture (or a band of CTU rows) per call:
 *   search_intra_rough (search_intra.c:986-1062): every n x n block of every rough level is predicted with the level's
 *     k candidate modes and scored cost = min(SATD, 2 SAD) (get_cost, search_intra.c:133-192);
 *   mode selection (search_intra.c:1064-1105): the n_best cheapest candidates per block, best first, ties in
 *     candidate order;
 *   uvg_quantize_lcu_residual (transform.c:1487-1600) for every TU of the configured sizes: intra prediction with the
 *     TU's mode, quantize_residual (quant-generic.c:460-612), pixels_calc_ssd, fast_coeff_cost.
 * Only the source plane goes to the device; predictions, residuals, reconstructions stay in HBM. Reference samples
 * are the SOURCE picture's neighbouring samples (open loop; border replication outside the picture).
 *
 * Addressing (all arrays): CTU-major, index = (ctu - ctu_begin) * per_ctu + j; CTUs in raster order, 64x64, the
 * picture padded to whole CTUs by border replication. A rough level has (64/n)^2 blocks per CTU, j = raster position
 * in the CTU. TU j of a CTU sits at raster position j % (64/n)^2 and is visit r = j / (64/n)^2 of that position
 * (the encoder's RDO revisits positions): it is predicted with candidate min(r, n_best-1) of the covering block of
 * the smallest rough level that is at least as large as the TU (else the largest level). */
#define UVG_B200_FRAME_MAX_LEVELS 4
typedef struct uvg_b200_frame_level { int32_t n; int32_t k; int8_t modes[68]; } uvg_b200_frame_level;
typedef struct uvg_b200_frame_tu { int32_t n; int32_t per_ctu; } uvg_b200_frame_tu;
typedef struct uvg_b200_frame_cfg {
  int32_t width, height;            /* luma plane, pixels                                              */
  int32_t bitdepth;                 /* 8                                                               */
  int32_t n_levels;                 /* rough-search levels, n in {8, 16, 32}                           */
  uvg_b200_frame_level level[UVG_B200_FRAME_MAX_LEVELS];
  int32_t n_tus;                    /* TU sizes of the residual pass, n in {4, 8, 16, 32}              */
  uvg_b200_frame_tu tu[UVG_B200_FRAME_MAX_LEVELS];
  int32_t n_best;                   /* candidates kept per rough block, 1..4                           */
  uvg_b200_quant_cfg quant;         /* as for uvg_b200_quantize_residual_dev                           */
  uint64_t cost_weights;            /* fast_coeff_cost weights                                         */
} uvg_b200_frame_cfg;

/* Results of one pass: HOST pointers into pinned buffers owned by the frame object, valid until the next submit.
 * Coefficients come back compact: tu_sig[i] is a bit array over the dense [n_tus][n*n] coefficient array of TU size
 * i (coefficient g = bit (g & 31) of word g >> 5), `levels` holds the non-zero coefficients of all TUs - a TU's are
 * contiguous and in index order, starting at tu_level_off[i][t]; the order of the TUs inside `levels` is unspecified
 * (it can differ between two passes over the same picture). uvg_b200_frame_expand_coeffs rebuilds the dense array. tu_coeff_dense is set only with
 * UVG_B200_FRAME_DENSE_COEFF. */
typedef struct uvg_b200_frame_result {
  int32_t ctu_begin, ctu_end;
  const uint32_t *rough_cost[UVG_B200_FRAME_MAX_LEVELS];   /* [blocks][k]; only with UVG_B200_FRAME_ALL_COSTS */
  const uint8_t *rough_best[UVG_B200_FRAME_MAX_LEVELS];    /* [blocks][n_best] mode values, best first   */
  const uint32_t *rough_best_cost[UVG_B200_FRAME_MAX_LEVELS]; /* [blocks][n_best] their costs (what search_intra_rough returns: modes_out[i].cost) */
  const uint8_t *tu_mode[UVG_B200_FRAME_MAX_LEVELS];       /* [tus] the mode each TU was predicted with  */
  const uint8_t *tu_has_coeffs[UVG_B200_FRAME_MAX_LEVELS];
  const uint32_t *tu_ssd[UVG_B200_FRAME_MAX_LEVELS];
  const uint32_t *tu_cost[UVG_B200_FRAME_MAX_LEVELS];
  const uint32_t *tu_sig[UVG_B200_FRAME_MAX_LEVELS];
  const uint32_t *tu_level_off[UVG_B200_FRAME_MAX_LEVELS];
  const int16_t *tu_coeff_dense[UVG_B200_FRAME_MAX_LEVELS];
  const int16_t *levels;
  uint64_t n_levels;                /* non-zero coefficients of the pass                               */
  float device_ms;                  /* device time of the pass (H2D + kernels; CUDA events)            */
} uvg_b200_frame_result;

/* Device-side views for callers that keep working on the GPU (reconstruction for the next pictures, tests). */
typedef struct uvg_b200_frame_device {
  const void *plane;
  const uint32_t *rough_cost[UVG_B200_FRAME_MAX_LEVELS];
  const uint8_t *rough_best[UVG_B200_FRAME_MAX_LEVELS];
  const void *tu_pred[UVG_B200_FRAME_MAX_LEVELS], *tu_rec[UVG_B200_FRAME_MAX_LEVELS];
  const int16_t *tu_coeff[UVG_B200_FRAME_MAX_LEVELS];
  void *stream;
} uvg_b200_frame_device;

enum {
  UVG_B200_FRAME_NO_H2D = 1,        /* the source rows are already on the device (previous submit)     */
  UVG_B200_FRAME_NO_D2H = 2,        /* leave the results on the device                                 */
  UVG_B200_FRAME_DENSE_COEFF = 4,   /* additionally return the dense coefficient arrays                */
  UVG_B200_FRAME_ALL_COSTS = 8      /* additionally return the cost of every (block, mode), not only   */
                                    /* the n_best candidates' (the device-side copies always exist)    */
};

typedef struct uvg_b200_frame uvg_b200_frame;
UVG_B200_API uvg_b200_frame *uvg_b200_frame_create(const uvg_b200_frame_cfg *cfg);   /* NULL on failure */
UVG_B200_API void uvg_b200_frame_destroy(uvg_b200_frame *f);
/* Asynchronous: enqueues the H2D copy of the rows the band needs from `src` (host, `src_stride` bytes per row; pinned
 * memory makes the copy asynchronous), the kernels and the D2H copies on the object's own streams and returns.
 * CTU rows [ctu_row_begin, ctu_row_end); ctu_row_end <= 0 means "to the last row". One pass in flight per object;
 * several objects (pictures, or bands of one picture on several GPUs) run concurrently. */
UVG_B200_API int uvg_b200_frame_submit(uvg_b200_frame *f, const void *src, size_t src_stride,
                                       int ctu_row_begin, int ctu_row_end, int flags);
/* Blocks until the pass has finished and its results are in host memory. */
UVG_B200_API int uvg_b200_frame_wait(uvg_b200_frame *f, uvg_b200_frame_result *out);
UVG_B200_API int uvg_b200_frame_device_buffers(uvg_b200_frame *f, uvg_b200_frame_device *out);
/* Synchronous device -> host copy of `bytes` bytes from a pointer handed out by uvg_b200_frame_device_buffers. */
UVG_B200_API int uvg_b200_device_read(void *host_dst, const void *dev_src, size_t bytes);
UVG_B200_API void uvg_b200_frame_expand_coeffs(const uint32_t *sig, const uint32_t *level_off,
                                               const int16_t *levels, int coeffs_per_tu, size_t n_tus,
                                               int16_t *dense);

/* ---------------------------------------------------------------------------------- */
/* (5) Multi-GPU: the reference-picture exchange                       */
/* ---------------------------------------------------------------------------------- */
/* A picture is sharded across GPUs by CTU-row bands (one process per GPU). All-intra pictures need no communication.
 * For inter pictures every GPU must hold the whole reconstructed reference picture before the next picture's motion
 * search while it only reconstructs its own band: one broadcast per band, issued as ONE NCCL group over NVLink, in
 * place on the picture buffer. NCCL is bound at run time (libnccl.so.2); nothing of NCCL appears in this ABI.
 * Bootstrap: rank 0 calls uvg_b200_comm_unique_id and hands the 128 bytes to every rank through the host
 * application's own channel; every rank then calls uvg_b200_comm_create (collective). */
#define UVG_B200_COMM_ID_BYTES 128
typedef struct uvg_b200_comm uvg_b200_comm;
UVG_B200_API void uvg_b200_band_rows(int rank, int world, int ctu_rows, int *lo, int *hi);
UVG_B200_API int uvg_b200_comm_unique_id(void *id_out /* UVG_B200_COMM_ID_BYTES */);
UVG_B200_API uvg_b200_comm *uvg_b200_comm_create(const void *id, int rank, int world);   /* NULL on failure */
UVG_B200_API void uvg_b200_comm_destroy(uvg_b200_comm *c);
UVG_B200_API int uvg_b200_comm_nccl_version(void);                                       /* 0: NCCL not available */
/* plane: DEVICE pointer, `stride_bytes` per pixel row, `height` pixel rows; band r = CTU rows uvg_b200_band_rows(r).
 * Asynchronous on `stream`. */
UVG_B200_API int uvg_b200_refpic_exchange(uvg_b200_comm *c, void *plane, size_t stride_bytes, int height,
                                          void *stream);

/* Pinned host memory for callers that want the fast H2D path (optional). */
UVG_B200_API void *uvg_b200_host_alloc(size_t bytes);
UVG_B200_API void uvg_b200_host_free(void *p);

/* ---------------------------------------------------------------------------------- */
/* (1) Drop-in leaf functions (host pointers). `_8` = uint8_t pixels (UVG_BIT_DEPTH 8), */
/*     `_16` = uint16_t pixels (UVG_BIT_DEPTH 10); bitdepth for the _16 set is fixed by */
/*     uvg_b200_set_dropin_bitdepth (default 10).                                       */
/* ---------------------------------------------------------------------------------- */
UVG_B200_API void uvg_b200_set_dropin_bitdepth(int bitdepth);

#define UVG_B200_DECL_PICTURE(SUF, PIX)                                                        \
  UVG_B200_API unsigned uvg_b200_reg_sad_##SUF(const PIX *d1, const PIX *d2, int w, int h,      \
                                               unsigned stride1, unsigned stride2);            \
  UVG_B200_API unsigned uvg_b200_sad_4x4_##SUF(const PIX *a, const PIX *b);                     \
  UVG_B200_API unsigned uvg_b200_sad_8x8_##SUF(const PIX *a, const PIX *b);                     \
  UVG_B200_API unsigned uvg_b200_sad_16x16_##SUF(const PIX *a, const PIX *b);                   \
  UVG_B200_API unsigned uvg_b200_sad_32x32_##SUF(const PIX *a, const PIX *b);                   \
  UVG_B200_API unsigned uvg_b200_sad_64x64_##SUF(const PIX *a, const PIX *b);                   \
  UVG_B200_API unsigned uvg_b200_satd_4x4_##SUF(const PIX *a, const PIX *b);                    \
  UVG_B200_API unsigned uvg_b200_satd_8x8_##SUF(const PIX *a, const PIX *b);                    \
  UVG_B200_API unsigned uvg_b200_satd_16x16_##SUF(const PIX *a, const PIX *b);                  \
  UVG_B200_API unsigned uvg_b200_satd_32x32_##SUF(const PIX *a, const PIX *b);                  \
  UVG_B200_API unsigned uvg_b200_satd_64x64_##SUF(const PIX *a, const PIX *b);                  \
  /* preds = pred_buffer: candidate k at preds + k*1024 pixels (strategies-picture.h:48) */    \
  UVG_B200_API void uvg_b200_sad_4x4_dual_##SUF(const PIX *preds, const PIX *orig, unsigned n, unsigned *out);   \
  UVG_B200_API void uvg_b200_sad_8x8_dual_##SUF(const PIX *preds, const PIX *orig, unsigned n, unsigned *out);   \
  UVG_B200_API void uvg_b200_sad_16x16_dual_##SUF(const PIX *preds, const PIX *orig, unsigned n, unsigned *out); \
  UVG_B200_API void uvg_b200_sad_32x32_dual_##SUF(const PIX *preds, const PIX *orig, unsigned n, unsigned *out); \
  UVG_B200_API void uvg_b200_sad_64x64_dual_##SUF(const PIX *preds, const PIX *orig, unsigned n, unsigned *out); \
  UVG_B200_API void uvg_b200_satd_64x64_dual_##SUF(const PIX *preds, const PIX *orig, unsigned n, unsigned *out);\
  UVG_B200_API void uvg_b200_satd_4x4_dual_##SUF(const PIX *preds, const PIX *orig, unsigned n, unsigned *out);  \
  UVG_B200_API void uvg_b200_satd_8x8_dual_##SUF(const PIX *preds, const PIX *orig, unsigned n, unsigned *out);  \
  UVG_B200_API void uvg_b200_satd_16x16_dual_##SUF(const PIX *preds, const PIX *orig, unsigned n, unsigned *out);\
  UVG_B200_API void uvg_b200_satd_32x32_dual_##SUF(const PIX *preds, const PIX *orig, unsigned n, unsigned *out);\
  UVG_B200_API unsigned uvg_b200_satd_any_size_##SUF(int w, int h, const PIX *b1, int s1,       \
                                                     const PIX *b2, int s2);                   \
  UVG_B200_API unsigned uvg_b200_satd_any_size_vtm_##SUF(int w, int h, const PIX *b1, int s1,   \
                                                         const PIX *b2, int s2);               \
  UVG_B200_API void uvg_b200_satd_any_size_quad_##SUF(int w, int h, const PIX **preds,          \
                                                      int stride, const PIX *orig,             \
                                                      int orig_stride, unsigned num_modes,     \
                                                      unsigned *costs_out, int8_t *valid);     \
  UVG_B200_API unsigned uvg_b200_pixels_calc_ssd_##SUF(const PIX *ref, const PIX *rec,          \
                                                       int ref_stride, int rec_stride, int w,  \
                                                       int h);                                 \
  UVG_B200_API uint32_t uvg_b200_ver_sad_##SUF(const PIX *pic, const PIX *ref_row, int32_t w,   \
                                               int32_t h, uint32_t pic_stride);                \
  UVG_B200_API uint32_t uvg_b200_hor_sad_##SUF(const PIX *pic, const PIX *ref, int32_t w,       \
                                               int32_t h, uint32_t pic_stride,                 \
                                               uint32_t ref_stride, uint32_t left,             \
                                               uint32_t right);                                \
  UVG_B200_API void uvg_b200_generate_residual_##SUF(const PIX *ref, const PIX *pred,           \
                                                     int16_t *residual, int w, int h,          \
                                                     int ref_stride, int pred_stride);         \
  /* optimized_sad_func_ptr_t (optimized_sad_func_ptr_t.h:13): width-specialised SADs */       \
  typedef uint32_t (*uvg_b200_optimized_sad_##SUF)(const PIX *, const PIX *, int32_t height,    \
                                                   uint32_t stride1, uint32_t stride2);        \
  UVG_B200_API uvg_b200_optimized_sad_##SUF uvg_b200_get_optimized_sad_##SUF(int32_t width);

UVG_B200_DECL_PICTURE(8, uint8_t)
UVG_B200_DECL_PICTURE(16, uint16_t)

/* strategies-intra.h:46-80 (row 8(f)-2): the three leaves intra_predict_regular dispatches to
 * (intra.c:729, :733, :751), with cu_loc_t flattened to the w x h of the predicted plane block
 * (cu_loc->width/height for luma, chroma_width/chroma_height otherwise). `above` / `left` / `top` are
 * the rows of uvg_intra_ref (intra.h:47-50: 2*128+3+33*3 samples each, index 0 = corner). */
#define UVG_B200_DECL_INTRA(SUF, PIX)                                                           \
  UVG_B200_API void uvg_b200_angular_pred_##SUF(int w, int h, int intra_mode, int channel_type, \
                                                const PIX *above, const PIX *left, PIX *dst,    \
                                                int multi_ref_idx, int isp_mode);               \
  UVG_B200_API void uvg_b200_intra_pred_planar_##SUF(int w, int h, const PIX *top,              \
                                                     const PIX *left, PIX *dst);                \
  UVG_B200_API void uvg_b200_pdpc_planar_dc_##SUF(int mode, int w, int h, const PIX *top,       \
                                                  const PIX *left, PIX *dst);
UVG_B200_DECL_INTRA(8, uint8_t)
UVG_B200_DECL_INTRA(16, uint16_t)

/* strategies-ipol.h:105-108 (row 8(f)-1): the four MC sample-interpolation leaves; the unused
 * encoder_control_t / hor_flag / ver_flag arguments are dropped and mv[2] is passed as two ints
 * (only mv & 15 / mv & 31 is used, like the reference). src = the block's top-left integer sample. */
#define UVG_B200_DECL_IPOL(SUF, PIX)                                                            \
  UVG_B200_API void uvg_b200_sample_quarterpel_luma_##SUF(const PIX *src, int src_stride, int w, \
                                                          int h, PIX *dst, int dst_stride,      \
                                                          int mvx, int mvy);                    \
  UVG_B200_API void uvg_b200_sample_quarterpel_luma_hi_##SUF(const PIX *src, int src_stride,     \
                                                             int w, int h, int16_t *dst,        \
                                                             int dst_stride, int mvx, int mvy); \
  UVG_B200_API void uvg_b200_sample_octpel_chroma_##SUF(const PIX *src, int src_stride, int w,   \
                                                        int h, PIX *dst, int dst_stride,        \
                                                        int mvx, int mvy);                      \
  UVG_B200_API void uvg_b200_sample_octpel_chroma_hi_##SUF(const PIX *src, int src_stride,       \
                                                           int w, int h, int16_t *dst,          \
                                                           int dst_stride, int mvx, int mvy);
UVG_B200_DECL_IPOL(8, uint8_t)
UVG_B200_DECL_IPOL(16, uint16_t)
/* strategies-ipol.h:62-64, :142-145: the four FME block filters filter_hpel_blocks_hor_ver_luma (step
 * 0), filter_hpel_blocks_diag_luma (1), filter_qpel_blocks_hor_ver_luma (2), filter_qpel_blocks_diag_luma
 * (3). filtered = the reference's `uvg_pixel filtered[4][LCU_LUMA_SIZE]` (block_pitch = LCU_LUMA_SIZE,
 * filtered_stride = LCU_WIDTH); off_x / off_y = the best half-pel offset (steps 2, 3). The shared
 * intermediate buffers of the reference are not used. */
UVG_B200_API void uvg_b200_fme_filter_8(int step, const uint8_t *src, int src_stride, int w, int h,
                                        uint8_t *filtered, int filtered_stride, int block_pitch,
                                        int off_x, int off_y);
UVG_B200_API void uvg_b200_fme_filter_16(int step, const uint16_t *src, int src_stride, int w, int h,
                                         uint16_t *filtered, int filtered_stride, int block_pitch,
                                         int off_x, int off_y);

/* crc32c_4x4 / crc32c_8x8 with the reference's signature; bipred_average per plane (the glue splits the
 * composite's lcu_t / yuv_t arguments into planes exactly like bipred_average_generic does). */
#define UVG_B200_DECL_MISC(SUF, PIX)                                                            \
  UVG_B200_API uint32_t uvg_b200_crc32c_4x4_##SUF(const PIX *buf, uint32_t pic_stride);         \
  UVG_B200_API uint32_t uvg_b200_crc32c_8x8_##SUF(const PIX *buf, uint32_t pic_stride);         \
  UVG_B200_API void uvg_b200_bipred_average_##SUF(PIX *dst, int dst_stride, const void *src0,   \
                                                  int src0_is_im, const void *src1,             \
                                                  int src1_is_im, int w, int h);
UVG_B200_DECL_MISC(8, uint8_t)
UVG_B200_DECL_MISC(16, uint16_t)
UVG_B200_API double uvg_b200_pixel_var_8(const uint8_t *buf, uint32_t len);
UVG_B200_API double uvg_b200_pixel_var_16(const uint16_t *buf, uint32_t len);

/* strategies-dct.h:44 dct_func (the typedef's unsigned return is never used; void here) */
UVG_B200_API void uvg_b200_dct_4x4(int8_t bitdepth, const int16_t *in, int16_t *out);
UVG_B200_API void uvg_b200_dct_8x8(int8_t bitdepth, const int16_t *in, int16_t *out);
UVG_B200_API void uvg_b200_dct_16x16(int8_t bitdepth, const int16_t *in, int16_t *out);
UVG_B200_API void uvg_b200_dct_32x32(int8_t bitdepth, const int16_t *in, int16_t *out);
UVG_B200_API void uvg_b200_idct_4x4(int8_t bitdepth, const int16_t *in, int16_t *out);
UVG_B200_API void uvg_b200_idct_8x8(int8_t bitdepth, const int16_t *in, int16_t *out);
UVG_B200_API void uvg_b200_idct_16x16(int8_t bitdepth, const int16_t *in, int16_t *out);
UVG_B200_API void uvg_b200_idct_32x32(int8_t bitdepth, const int16_t *in, int16_t *out);
UVG_B200_API void uvg_b200_fast_forward_dst_4x4(int8_t bitdepth, const int16_t *in, int16_t *out);
UVG_B200_API void uvg_b200_fast_inverse_dst_4x4(int8_t bitdepth, const int16_t *in, int16_t *out);
/* mts_dct_func (strategies-dct.h:63-71) with cu_info_t flattened to the five fields the
 * reference reads (dct-generic.c:2501-2557): tu->type, tr_idx, lfnst_idx, cr_lfnst_idx,
 * intra.isp_mode. */
UVG_B200_API void uvg_b200_mts_dct(int8_t bitdepth, int color, int cu_type, int tr_idx,
                                   int lfnst_idx, int cr_lfnst_idx, int isp_mode, int8_t w,
                                   int8_t h, const int16_t *in, int16_t *out, int8_t mts_type);
UVG_B200_API void uvg_b200_mts_idct(int8_t bitdepth, int color, int cu_type, int tr_idx,
                                    int lfnst_idx, int cr_lfnst_idx, int isp_mode, int8_t w,
                                    int8_t h, const int16_t *in, int16_t *out, int8_t mts_type);
UVG_B200_API void uvg_b200_get_tr_type(int w, int h, int color, int cu_type, int tr_idx,
                                       int lfnst_idx, int cr_lfnst_idx, int isp_mode,
                                       int mts_type, int *hor, int *ver);

/* strategies-quant.h:48-91 with encoder_state_t flattened into uvg_b200_quant_cfg. */
UVG_B200_API void uvg_b200_quant(const uvg_b200_quant_cfg *cfg, const int16_t *coef,
                                 int16_t *q_coef, int32_t w, int32_t h, int color,
                                 int8_t scan_idx, int8_t block_type, int8_t transform_skip,
                                 uint8_t lfnst_idx);
UVG_B200_API void uvg_b200_dequant(const uvg_b200_quant_cfg *cfg, const int16_t *q_coef,
                                   int16_t *coef, int32_t w, int32_t h, int color,
                                   int8_t block_type, int8_t transform_skip);
UVG_B200_API int uvg_b200_quantize_residual_8(const uvg_b200_quant_cfg *cfg, int cu_type,
                                              int tr_idx, int w, int h, int color,
                                              int use_trskip, int in_stride, int out_stride,
                                              const uint8_t *ref_in, const uint8_t *pred_in,
                                              uint8_t *rec_out, int16_t *coeff_out,
                                              int early_skip);
UVG_B200_API int uvg_b200_quantize_residual_16(const uvg_b200_quant_cfg *cfg, int cu_type,
                                               int tr_idx, int w, int h, int color,
                                               int use_trskip, int in_stride, int out_stride,
                                               const uint16_t *ref_in, const uint16_t *pred_in,
                                               uint16_t *rec_out, int16_t *coeff_out,
                                               int early_skip);
/* quant_cbcr_residual (strategies-quant.h:62-80): cur_cu->joint_cb_cr and state->frame->jccr_sign
 * travel as plain ints; returns joint_cb_cr when coefficients survive, else 0. */
UVG_B200_API int uvg_b200_quant_cbcr_residual_8(const uvg_b200_quant_cfg *cfg, int cu_type, int tr_idx,
                                                int w, int h, int joint_cb_cr, int jccr_sign,
                                                int in_stride, int out_stride,
                                                const uint8_t *u_ref, const uint8_t *v_ref,
                                                const uint8_t *u_pred, const uint8_t *v_pred,
                                                uint8_t *u_rec, uint8_t *v_rec, int16_t *coeff_out,
                                                int early_skip);
UVG_B200_API int uvg_b200_quant_cbcr_residual_16(const uvg_b200_quant_cfg *cfg, int cu_type, int tr_idx,
                                                 int w, int h, int joint_cb_cr, int jccr_sign,
                                                 int in_stride, int out_stride,
                                                 const uint16_t *u_ref, const uint16_t *v_ref,
                                                 const uint16_t *u_pred, const uint16_t *v_pred,
                                                 uint16_t *u_rec, uint16_t *v_rec, int16_t *coeff_out,
                                                 int early_skip);
UVG_B200_API uint32_t uvg_b200_coeff_abs_sum(const int16_t *coeffs, size_t length);
UVG_B200_API uint32_t uvg_b200_fast_coeff_cost(const int16_t *coeff, int32_t w, int32_t h,
                                               uint64_t weights);

#ifdef __cplusplus
}
#endif
#endif /* UVG266_B200_H_ */
