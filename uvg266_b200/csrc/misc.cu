// misc.cu - the two remaining non-distortion entries of the picture strategy table:
//   crc32c_4x4 / crc32c_8x8  (picture-generic.c:1371-1443): the IBC block hash, CRC-32C (reflected
//       polynomial 0x82F63B78) over the block's pixel bytes, row by row;
//   bipred_average          (picture-generic.c:1133-1247): rounded average of two predictions, each
//       either pixels or the 14-bit intermediates of the "_hi" interpolation (ipol.cu).
// pixel_var (:1334-1357) accumulates squared deviations in double in index order; a parallel
// reduction cannot reproduce its rounding, so it stays with the reference's implementation.
#include "common.cuh"

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
