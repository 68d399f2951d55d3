// search.cu - integer motion-search SAD over a full (2r+1)^2 window, 8-bit fast path.
//
// Call site it batches: search_inter.c:204-247 check_mv_cost -> image.c:438 uvg_image_calc_sad
// (-> reg_sad, or ver_sad / hor_sad / cor_sad when the candidate leaves the frame). One CTA per
// current block: the (w+2r) x (h+2r) reference window is staged ONCE in shared memory with
// clamped coordinates (== the reference's edge replication), the current block too; HBM traffic
// per block is the unique bytes of the window instead of (2r+1)^2 * w*h.
//
// Each thread owns four horizontally adjacent candidates (dx = 4j .. 4j+3, same dy): one aligned
// 32-bit shared-memory word of the window feeds all four through funnel shifts, and every
// VABSDIFF4.U8.ACC consumes 4 pixel pairs -> 2.25 instructions per 4 pixel pairs.
// Lanes of a warp hold consecutive j, i.e. consecutive words of a window row: no bank conflicts
// inside a row; the row pitch is chosen so that the second row a warp may touch starts on the
// bank after the first row's last group.
#include "common.cuh"
#include <mutex>

namespace uvgb {

__global__ void __launch_bounds__(256)
sad_search_u8_kernel(const uint8_t *__restrict__ cur, const uint8_t *__restrict__ ref, int pic_w, int pic_h,
                     int stride, const int32_t *__restrict__ cur_x, const int32_t *__restrict__ cur_y,
                     const int32_t *__restrict__ mv_x, const int32_t *__restrict__ mv_y, int range,
                     int w, int h, int pitch_words, uint32_t *__restrict__ out)
{
  extern __shared__ __align__(16) uint32_t smem_w[];
  const int wh = h + 2 * range;
  uint32_t *s_ref = smem_w;                               // wh rows of pitch_words words
  uint32_t *s_cur = smem_w + (size_t)pitch_words * wh;    // h rows of w/4 words
  const size_t j = blockIdx.x;
  const int cx = cur_x[j], cy = cur_y[j];
  const int ox = cx + mv_x[j] - range, oy = cy + mv_y[j] - range;
  const int pitch = pitch_words * 4;
  uint8_t *s_ref_b = reinterpret_cast<uint8_t *>(s_ref);
  for (int t = threadIdx.x; t < pitch * wh; t += blockDim.x) {
    const int y = t / pitch, x = t - y * pitch;
    const int rx = clip3(0, pic_w - 1, ox + x), ry = clip3(0, pic_h - 1, oy + y);
    s_ref_b[t] = __ldg(ref + (size_t)ry * stride + rx);
  }
  uint8_t *s_cur_b = reinterpret_cast<uint8_t *>(s_cur);
  for (int t = threadIdx.x; t < w * h; t += blockDim.x) {
    const int y = t / w, x = t - y * w;
    s_cur_b[t] = __ldg(cur + (size_t)(cy + y) * stride + cx + x);
  }
  __syncthreads();
  const int side = 2 * range + 1, g4 = (side + 3) >> 2, ww = w >> 2;
  const int n_groups = g4 * side;
  for (int c = threadIdx.x; c < n_groups; c += blockDim.x) {
    const int dy = c / g4, jg = c - dy * g4;
    unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int y = 0; y < h; ++y) {
      const uint32_t *r = s_ref + (dy + y) * pitch_words + jg;
      const uint32_t *p = s_cur + y * ww;
      uint32_t lo = r[0];
#pragma unroll 4
      for (int i = 0; i < ww; ++i) {
        const uint32_t hi = r[i + 1];
        const uint32_t pc = p[i];
        a0 = __vsadu4(lo, pc) + a0;
        a1 = __vsadu4(__funnelshift_r(lo, hi, 8), pc) + a1;
        a2 = __vsadu4(__funnelshift_r(lo, hi, 16), pc) + a2;
        a3 = __vsadu4(__funnelshift_r(lo, hi, 24), pc) + a3;
        lo = hi;
      }
    }
    uint32_t *o = out + j * (size_t)(side * side) + (size_t)dy * side + 4 * jg;
    const int dx0 = 4 * jg;
    o[0] = a0;
    if (dx0 + 1 < side) o[1] = a1;
    if (dx0 + 2 < side) o[2] = a2;
    if (dx0 + 3 < side) o[3] = a3;
  }
}

// -1: not covered (non-8-bit, width not a multiple of 4, window too large)
int launch_sad_search_u8(const uint8_t *cur, const uint8_t *ref, int pic_w, int pic_h, int stride,
                         const int32_t *cur_x, const int32_t *cur_y, const int32_t *mv_x, const int32_t *mv_y,
                         int range, int w, int h, uint32_t *out, size_t n_blocks, cudaStream_t st)
{
  if (w % 4) return -1;
  const int side = 2 * range + 1, g4 = (side + 3) / 4;
  // a thread reads words jg .. jg + w/4 of a row: the row needs g4 + w/4 words
  int pitch_words = g4 + w / 4;
  if (g4 <= 32) while ((pitch_words & 31) != (g4 & 31)) ++pitch_words;   // see header comment
  const size_t smem = ((size_t)pitch_words * (h + 2 * range) + (size_t)(w / 4) * h) * 4;
  if (smem > 227 * 1024) return -1;
  // callable from any number of threads: raise the limit once, to the 227 KB cap
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(sad_search_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  UVGB_TRY(attr_err);
  sad_search_u8_kernel<<<(unsigned)n_blocks, 256, smem, st>>>(cur, ref, pic_w, pic_h, stride, cur_x, cur_y, mv_x,
                                                              mv_y, range, w, h, pitch_words, out);
  UVGB_LAUNCH_CHECK();
  return 0;
}

}  // namespace uvgb
