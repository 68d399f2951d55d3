// misc.cu - the two remaining non-distortion entries of the picture strategy table:
//   crc32c_4x4 / crc32c_8x8  (picture-generic.c:1371-1443): the IBC block hash, CRC-32C (reflected
//       polynomial 0x82F63B78) over the block's pixel bytes, row by row;
//   bipred_average          (picture-generic.c:1133-1247): rounded average of two predictions, each
//       either pixels or the 14-bit intermediates of the "_hi" interpolation (ipol.cu).
//   pixel_var               (picture-generic.c:1334-1357): variance of a pixel array in double. The mean is a sum of
//       integers (exact in any order); the squared deviations are accumulated in INDEX ORDER, one rounding per
//       addition, which no parallel reduction reproduces - so one thread walks the array in order while its warp
//       streams the pixels through shared memory, and (v - mean)^2 comes from a per-array table of the <= 2^bitdepth
//       distinct values (computed with the same two roundings as the reference's `tmp = v - mean; tmp * tmp`).
#include "common.cuh"
#include <mutex>
#include <string.h>

namespace uvgb {

__constant__ uint32_t c_crc32c[256];

static int ensure_crc_table()
{
  static bool done = false;
  if (done) return 0;
  uint32_t t[256];
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int b = 0; b < 8; ++b) c = (c >> 1) ^ ((c & 1) ? 0x82F63B78u : 0u);
    t[i] = c;
  }
  UVGB_TRY(cudaMemcpyToSymbol(c_crc32c, t, sizeof(t)));
  done = true;
  return 0;
}

// one thread per block; the encoder hashes every 8x8 / 4x4 block position of the source picture
template <typename Pix>
__global__ void crc32c_kernel(Blocks A, int n, uint32_t *__restrict__ out, size_t n_blocks)
{
  const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  const Pix *p = static_cast<const Pix *>(A.base) + A.offset(b);
  uint32_t crc = 0xFFFFFFFFu;
  for (int y = 0; y < n; ++y)
    for (int x = 0; x < n; ++x) {
      const uint32_t v = p[(size_t)y * A.stride + x];
      crc = (crc >> 8) ^ c_crc32c[(crc ^ v) & 0xFF];
      if (sizeof(Pix) == 2) crc = (crc >> 8) ^ c_crc32c[(crc ^ (v >> 8)) & 0xFF];
    }
  out[b] = crc ^ 0xFFFFFFFFu;
}

// dst[b][y][x] = clip((s0 + s1 + offset) >> (15 - bitdepth)); sources dense [n_blocks][w*h]
template <typename Pix>
__global__ void bipred_kernel(Blocks D, int w, int h, const void *__restrict__ s0, int im0, const void *__restrict__ s1,
                              int im1, int bitdepth, size_t total)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t b = i / (size_t)(w * h);
  const int r = (int)(i - b * (size_t)(w * h)), y = r / w, x = r - y * w;
  const int up = 14 - bitdepth, shift = 15 - bitdepth;
  const int a = im0 ? (int)static_cast<const int16_t *>(s0)[i] : (int)(int16_t)(static_cast<const Pix *>(s0)[i] << up);
  const int c = im1 ? (int)static_cast<const int16_t *>(s1)[i] : (int)(int16_t)(static_cast<const Pix *>(s1)[i] << up);
  Pix *d = const_cast<Pix *>(static_cast<const Pix *>(D.base)) + D.offset(b) + (size_t)y * D.stride + x;
  *d = (Pix)min(max((a + c + (1 << (shift - 1))) >> shift, 0), (1 << bitdepth) - 1);
}

// One CTA per array. out[a] = pixel_var(arrays + a * pitch, len).
template <typename Pix>
__global__ void __launch_bounds__(256)
pixel_var_kernel(const Pix *__restrict__ arrays, size_t pitch, uint32_t len, int n_values, double *__restrict__ out)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *s_sq = reinterpret_cast<double *>(smem_raw);                    // [n_values]
  Pix *s_chunk = reinterpret_cast<Pix *>(s_sq + n_values);                // [2][CH]
  constexpr int CH = 4096;
  __shared__ unsigned long long s_sum;
  const Pix *p = arrays + (size_t)blockIdx.x * pitch;
  if (threadIdx.x == 0) s_sum = 0;
  __syncthreads();
  unsigned long long part = 0;
  for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) part += p[i];
  atomicAdd(&s_sum, part);
  __syncthreads();
  const double mean = __ddiv_rn((double)s_sum, (double)len);               // sum / (double)len, :1346
  for (int v = threadIdx.x; v < n_values; v += blockDim.x) {
    const double tmp = __dsub_rn((double)v, mean);                         // (double)arr[i] - arr_mean
    s_sq[v] = __dmul_rn(tmp, tmp);
  }
  double var = 0.0;
  for (uint32_t base = 0, buf = 0; base < len; base += CH, buf ^= 1) {
    const uint32_t cnt = min((uint32_t)CH, len - base);
    Pix *c = s_chunk + buf * CH;
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) c[i] = p[base + i];
    __syncthreads();                                                       // also orders the table before its first use
    if (threadIdx.x == 0)
      for (uint32_t i = 0; i < cnt; ++i) var = __dadd_rn(var, s_sq[c[i]]);  // var += tmp * tmp, in index order
    // double buffering: the other threads go on to stage the next chunk while thread 0 is still adding
  }
  if (threadIdx.x == 0) out[blockIdx.x] = __ddiv_rn(var, (double)len);      // var /= len, :1354
}

}  // namespace uvgb

using namespace uvgb;

extern "C" int uvg_b200_crc32c_dev(int bitdepth, int n, const uvg_b200_blocks *blocks, uint32_t *out, size_t n_blocks,
                                   void *stream)
{
  if (ensure_init() || ensure_crc_table()) return 1;
  if (n != 4 && n != 8) { set_error("crc32c: n must be 4 or 8 (got %d)", n); return 1; }
  if (n_blocks == 0) return 0;
  const Blocks A = to_blocks(blocks);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((n_blocks + 127) / 128);
  if (bitdepth == 8) crc32c_kernel<uint8_t><<<grid, 128, 0, st>>>(A, n, out, n_blocks);
  else crc32c_kernel<uint16_t><<<grid, 128, 0, st>>>(A, n, out, n_blocks);
  UVGB_LAUNCH_CHECK();
  return 0;
}

extern "C" int uvg_b200_bipred_average_dev(int bitdepth, int w, int h, const uvg_b200_blocks *dst, const void *src0,
                                           int src0_is_im, const void *src1, int src1_is_im, size_t n_blocks, void *stream)
{
  if (ensure_init()) return 1;
  if (bitdepth < 8 || bitdepth > 14 || w < 1 || h < 1) { set_error("bipred_average: bad bitdepth / size"); return 1; }
  if (n_blocks == 0) return 0;
  const Blocks D = to_blocks(dst);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t total = n_blocks * (size_t)w * h;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (bitdepth == 8) bipred_kernel<uint8_t><<<grid, 256, 0, st>>>(D, w, h, src0, src0_is_im, src1, src1_is_im, bitdepth, total);
  else bipred_kernel<uint16_t><<<grid, 256, 0, st>>>(D, w, h, src0, src0_is_im, src1, src1_is_im, bitdepth, total);
  UVGB_LAUNCH_CHECK();
  return 0;
}

// ---- drop-ins (host pointers, synchronous) ----------------------------------------------------------
namespace {
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
[[noreturn]] void die() { fprintf(stderr, "uvg266_b200: %s\n", uvg_b200_last_error()); abort(); }

template <typename Pix>
uint32_t crc_leaf(int n, const Pix *buf, uint32_t stride)
{
  if (ensure_init() || ensure_crc_table()) die();
  Staging &s = tls_staging();
  if (s.ensure(512, 1024)) die();
  Pix *hw = reinterpret_cast<Pix *>(s.h_pin);
  for (int y = 0; y < n; ++y) memcpy(hw + y * n, buf + (size_t)y * stride, n * sizeof(Pix));
  UVGB_MUST(cudaMemcpyAsync(s.d_buf, hw, (size_t)n * n * sizeof(Pix), cudaMemcpyHostToDevice, s.stream));
  uvg_b200_blocks bl = { s.d_buf, (uint32_t)n, 1, 0, (uint64_t)n * n };
  uint32_t *d_out = reinterpret_cast<uint32_t *>(s.d_buf + 512);
  if (uvg_b200_crc32c_dev(dropin_bitdepth(sizeof(Pix)), n, &bl, d_out, 1, s.stream)) die();
  uint32_t *h_out = reinterpret_cast<uint32_t *>(s.h_pin + 256);
  UVGB_MUST(cudaMemcpyAsync(h_out, d_out, 4, cudaMemcpyDeviceToHost, s.stream));
  UVGB_MUST(cudaStreamSynchronize(s.stream));
  return *h_out;
}

template <typename Pix>
void bipred_leaf(Pix *dst, int dst_stride, const void *s0, int im0, const void *s1, int im1, int w, int h)
{
  if (ensure_init()) die();
  if (w < 1 || h < 1 || w > 64 || h > 64) { fprintf(stderr, "uvg266_b200: bipred_average: block %dx%d unsupported\n", w, h); abort(); }
  Staging &s = tls_staging();
  const size_t e0 = im0 ? 2 : sizeof(Pix), e1 = im1 ? 2 : sizeof(Pix), cnt = (size_t)w * h;
  const size_t o1 = align_up(cnt * e0, 256), od = o1 + align_up(cnt * e1, 256);
  if (s.ensure(od + align_up(cnt * sizeof(Pix), 256), od + align_up(cnt * sizeof(Pix), 256))) die();
  memcpy(s.h_pin, s0, cnt * e0); memcpy(s.h_pin + o1, s1, cnt * e1);
  UVGB_MUST(cudaMemcpyAsync(s.d_buf, s.h_pin, od, cudaMemcpyHostToDevice, s.stream));
  uvg_b200_blocks bl = { s.d_buf + od, (uint32_t)w, 1, 0, (uint64_t)cnt };
  if (uvg_b200_bipred_average_dev(dropin_bitdepth(sizeof(Pix)), w, h, &bl, s.d_buf, im0, s.d_buf + o1, im1, 1, s.stream)) die();
  UVGB_MUST(cudaMemcpyAsync(s.h_pin + od, s.d_buf + od, cnt * sizeof(Pix), cudaMemcpyDeviceToHost, s.stream));
  UVGB_MUST(cudaStreamSynchronize(s.stream));
  const Pix *ho = reinterpret_cast<const Pix *>(s.h_pin + od);
  for (int y = 0; y < h; ++y) memcpy(dst + (size_t)y * dst_stride, ho + (size_t)y * w, (size_t)w * sizeof(Pix));
}
}  // namespace

#define UVGB_MISC_LEAVES(SUF, PIX)                                                                                 \
  extern "C" uint32_t uvg_b200_crc32c_4x4_##SUF(const PIX *buf, uint32_t pic_stride) { return crc_leaf<PIX>(4, buf, pic_stride); } \
  extern "C" uint32_t uvg_b200_crc32c_8x8_##SUF(const PIX *buf, uint32_t pic_stride) { return crc_leaf<PIX>(8, buf, pic_stride); } \
  extern "C" void uvg_b200_bipred_average_##SUF(PIX *dst, int dst_stride, const void *src0, int src0_is_im,        \
                                                const void *src1, int src1_is_im, int w, int h)                  \
  { bipred_leaf<PIX>(dst, dst_stride, src0, src0_is_im, src1, src1_is_im, w, h); }
UVGB_MISC_LEAVES(8, uint8_t)
UVGB_MISC_LEAVES(16, uint16_t)

// pixel_var of n_arrays arrays of `len` pixels, array a at arrays + a * pitch (pixels). One CTA per array.
extern "C" int uvg_b200_pixel_var_dev(int bitdepth, const void *arrays, size_t pitch, uint32_t len, double *out,
                                      size_t n_arrays, void *stream)
{
  if (ensure_init()) return 1;
  if (bitdepth < 8 || bitdepth > 12) { set_error("pixel_var: bitdepth %d unsupported", bitdepth); return 1; }
  if (n_arrays == 0) return 0;
  if (len == 0) { set_error("pixel_var: empty array"); return 1; }
  const int nv = 1 << bitdepth, pix = bitdepth == 8 ? 1 : 2;
  const size_t smem = (size_t)nv * sizeof(double) + 2 * 4096 * pix;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (bitdepth == 8) {
    pixel_var_kernel<uint8_t><<<(unsigned)n_arrays, 256, smem, st>>>(static_cast<const uint8_t *>(arrays), pitch, len, nv, out);
  } else {
    static std::once_flag once;
    static cudaError_t err = cudaSuccess;
    std::call_once(once, [] { err = cudaFuncSetAttribute(pixel_var_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); });
    UVGB_TRY(err);
    pixel_var_kernel<uint16_t><<<(unsigned)n_arrays, 256, smem, st>>>(static_cast<const uint16_t *>(arrays), pitch, len, nv, out);
  }
  UVGB_LAUNCH_CHECK();
  return 0;
}

template <typename Pix>
static double dropin_pixel_var(const Pix *buf, uint32_t len)
{
  if (ensure_init()) die();
  Staging &s = tls_staging();
  const size_t bytes = ((size_t)len * sizeof(Pix) + 255) / 256 * 256;
  if (s.ensure(16, bytes + 16)) die();
  UVGB_MUST(cudaMemcpyAsync(s.d_buf, buf, (size_t)len * sizeof(Pix), cudaMemcpyHostToDevice, s.stream));
  double *d_out = reinterpret_cast<double *>(s.d_buf + bytes);
  if (uvg_b200_pixel_var_dev(dropin_bitdepth(sizeof(Pix)), s.d_buf, len, len, d_out, 1, s.stream)) die();
  UVGB_MUST(cudaMemcpyAsync(s.h_pin, d_out, sizeof(double), cudaMemcpyDeviceToHost, s.stream));
  UVGB_MUST(cudaStreamSynchronize(s.stream));
  double r;
  memcpy(&r, s.h_pin, sizeof(r));
  return r;
}
extern "C" double uvg_b200_pixel_var_8(const uint8_t *buf, uint32_t len) { return dropin_pixel_var<uint8_t>(buf, len); }
extern "C" double uvg_b200_pixel_var_16(const uint16_t *buf, uint32_t len) { return dropin_pixel_var<uint16_t>(buf, len); }
