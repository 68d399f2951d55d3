// quant.cuh - parameter blocks shared by the quant / fused kernels (quant.cu,
// fused_small.cu, fused_mma.cu).
#pragma once
#include "transform.cuh"

namespace uvgb {

struct QParams {
  int w, h, area;
  // quant
  int q_bits, add, q_bits8, coeff_default, coeff_list, signhide, lfnst_idx, max_lfnst;
  const uint16_t *scan;
  // dequant
  int dq_shift, dq_scale, dq_add;
  // reconstruction
  int pixel_max, ssd_shift;
  // LFNST stage of the general kernel (0 = off): kernel matrix [16][tr_size] of the (set, index) pair on the device
  int lfnst_apply, lfnst_transpose, lfnst_sb, lfnst_zero_out, lfnst_whge3;
  const int8_t *lfnst_mat;
};


struct FusedArgs {
  Blocks ref, pred, rec;
  int16_t *coeff_out;
  uint8_t *has_coeffs;
  uint32_t *ssd;
  uint32_t *cost;
  unsigned w0, w1, w2, w3;   // fast_coeff_cost weights (Q8.8), rdo.c:440
  int use_trskip, early_skip, rec_is_pred;
  // joint Cb-Cr mode (uvg_quant_cbcr_residual_generic, quant-generic.c:241-442), general kernel only:
  // ref/pred/rec = the Cb plane, ref2/pred2/rec2 = the Cr plane, jccr = joint_cb_cr * (jccr_sign ? -1 : 1)
  Blocks ref2, pred2, rec2;
  int jccr;
};


// Thread-per-TU register kernel for square 4x4 / 8x8 TUs (fused_small.cu). Returns -1 when
// the shape/config is not covered (caller falls through to the general kernel).
int launch_fused_small(const FusedArgs &a, size_t n_tus, const QParams &q, int bitdepth, int type_hor,
                       int type_ver, cudaStream_t st);
// Tensor-core (mma.sync) kernel for square 16x16 / 32x32 TUs (fused_mma.cu); -1 = not covered.
int launch_fused_mma(const FusedArgs &a, size_t n_tus, const QParams &q, int bitdepth, int type_hor,
                     int type_ver, int skip_w, int skip_h, cudaStream_t st);

// INT8 tensor-core kernel for 8-bit square 8x8 / 16x16 / 32x32 TUs (fused_imma.cu); -1 = not covered.
int launch_fused_imma(const FusedArgs &a, size_t n_tus, const QParams &q, int bitdepth, int type_hor,
                      int type_ver, int skip_w, int skip_h, cudaStream_t st);

}  // namespace uvgb
