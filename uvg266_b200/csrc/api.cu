// api.cu - library state, host-side batched entry points and the drop-in leaf functions
// (reference strategy signatures, host pointers) of libuvg266_b200.
//
// Threading contract of the boundary: leaf functions are called
// concurrently from any number of encoder worker threads with no context argument, so
// every calling thread gets its own stream, pinned staging buffer and device scratch
// (thread_local Staging). There is no CPU fallback: a device error in a leaf function
// aborts the process with a message.
#include "common.cuh"
#include "transform.cuh"
#include <stdarg.h>
#include <string.h>
#include <mutex>

namespace uvgb {

int init_transform_tables();
int init_quant_tables();

std::atomic<uint64_t> g_launches{0};
static thread_local char t_err[512] = "";
static std::mutex g_init_mu;
static std::atomic<int> g_device{-1};
static int g_sms = 0;
static std::atomic<int> g_dropin_bd{10};

void set_error(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

void fatal(const char *what, cudaError_t e, const char *file, int line)
{
  fprintf(stderr, "uvg266_b200: fatal device error: %s: %s (%s:%d); there is no CPU fallback\n", what,
          cudaGetErrorString(e), file, line);
  abort();
}

static int do_init(int device)
{
  std::lock_guard<std::mutex> lk(g_init_mu);
  if (g_device.load() >= 0) {
    if (device >= 0 && device != g_device.load()) { set_error("already initialised on device %d", g_device.load()); return 1; }
    return 0;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device available (%s); uvg266_b200 has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return 1;
  }
  if (device < 0) device = 0;
  if (device >= n) { set_error("device %d out of range (%d devices)", device, n); return 1; }
  UVGB_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  UVGB_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return 1;
  }
  g_sms = prop.multiProcessorCount;
  if (init_transform_tables()) return 1;
  if (init_quant_tables()) return 1;
  g_device.store(device);
  return 0;
}

int ensure_init()
{
  if (g_device.load() >= 0) {
    // worker threads created after init still need the device selected
    static thread_local bool t_set = false;
    if (!t_set) { if (cudaSetDevice(g_device.load()) != cudaSuccess) { set_error("cudaSetDevice failed"); return 1; } t_set = true; }
    return 0;
  }
  return do_init(-1);
}

int sm_count() { return g_sms ? g_sms : 148; }

int Staging::ensure(size_t host_bytes, size_t dev_bytes)
{
  if (!stream) UVGB_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  if (host_bytes > h_cap) {
    if (h_pin) cudaFreeHost(h_pin);
    h_pin = nullptr; h_cap = 0;
    size_t cap = host_bytes < (1u << 16) ? (1u << 16) : host_bytes * 2;
    UVGB_TRY(cudaMallocHost(reinterpret_cast<void **>(&h_pin), cap));
    h_cap = cap;
  }
  if (dev_bytes > d_cap) {
    if (d_buf) { cudaStreamSynchronize(stream); cudaFree(d_buf); }
    d_buf = nullptr; d_cap = 0;
    size_t cap = dev_bytes < (1u << 16) ? (1u << 16) : dev_bytes * 2;
    UVGB_TRY(cudaMalloc(reinterpret_cast<void **>(&d_buf), cap));
    d_cap = cap;
  }
  return 0;
}

// Worker threads come and go (uvg266 open/close as a library): give back what the thread held. At process
// exit the driver may already be gone; every call then fails harmlessly and the error is dropped.
Staging::~Staging()
{
  if (!stream && !h_pin && !d_buf) return;
  if (cudaSetDevice(g_device.load()) != cudaSuccess) { cudaGetLastError(); return; }
  if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
  if (h_pin) cudaFreeHost(h_pin);
  if (d_buf) cudaFree(d_buf);
  cudaGetLastError();
}

Staging &tls_staging()
{
  static thread_local Staging s;
  return s;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Dense descriptor for n blocks of w*h packed back to back.
static uvg_b200_blocks dense_blocks(const void *base, int w, int h)
{
  uvg_b200_blocks b;
  b.base = base; b.stride = (uint32_t)w; b.blocks_per_row = 1; b.col_pitch = 0; b.row_pitch = (uint64_t)w * h;
  return b;
}

template <typename Pix>
static void pack_rows(Pix *dst, const Pix *src, int w, int h, size_t stride)
{
  for (int y = 0; y < h; ++y) memcpy(dst + (size_t)y * w, src + (size_t)y * stride, (size_t)w * sizeof(Pix));
}

// One distortion call on host blocks: a (w*h) against k candidates.
template <typename Pix>
static void dropin_distortion(int metric, int bitdepth, int w, int h, const Pix *a, size_t sa,
                              const Pix *const *cands, size_t sb, int k, unsigned *out)
{
  // A motion vector pointing entirely outside the frame leaves an empty middle part: the edge splitter calls
  // reg_sad / the width-specialised SADs with height (or width) 0 (image.c:404-416), whose loops do not run.
  if (w <= 0 || h <= 0) { for (int c = 0; c < k; ++c) out[c] = 0; return; }
  if (ensure_init()) { fprintf(stderr, "uvg266_b200: %s\n", t_err); abort(); }
  Staging &s = tls_staging();
  const size_t blk = (size_t)w * h * sizeof(Pix);
  const size_t in_bytes = align_up(blk * (1 + k), 16);
  const size_t total = in_bytes + 16 * sizeof(uint32_t);
  if (s.ensure(total, total)) { fprintf(stderr, "uvg266_b200: %s\n", t_err); abort(); }
  Pix *hp = reinterpret_cast<Pix *>(s.h_pin);
  pack_rows(hp, a, w, h, sa);
  for (int c = 0; c < k; ++c) pack_rows(hp + (size_t)(1 + c) * w * h, cands[c], w, h, sb);
  UVGB_MUST(cudaMemcpyAsync(s.d_buf, s.h_pin, in_bytes, cudaMemcpyHostToDevice, s.stream));
  const uvg_b200_blocks A = dense_blocks(s.d_buf, w, h);
  const uvg_b200_blocks B = dense_blocks(s.d_buf + blk, w, h);
  uint32_t *d_out = reinterpret_cast<uint32_t *>(s.d_buf + in_bytes);
  if (uvg_b200_distortion_dev(metric, bitdepth, w, h, &A, &B, d_out, 1, k, s.stream)) {
    fprintf(stderr, "uvg266_b200: %s\n", t_err); abort();
  }
  uint32_t *h_out = reinterpret_cast<uint32_t *>(s.h_pin + in_bytes);
  UVGB_MUST(cudaMemcpyAsync(h_out, d_out, (size_t)k * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
  UVGB_MUST(cudaStreamSynchronize(s.stream));
  for (int c = 0; c < k; ++c) out[c] = h_out[c];
}

template <typename Pix>
static unsigned dropin_pair(int metric, int bitdepth, int w, int h, const Pix *a, size_t sa, const Pix *b, size_t sb)
{
  unsigned r = 0;
  const Pix *c[1] = { b };
  dropin_distortion<Pix>(metric, bitdepth, w, h, a, sa, c, sb, 1, &r);
  return r;
}

int dropin_bitdepth(size_t pixel_size) { return pixel_size == 1 ? 8 : g_dropin_bd.load(); }
template <typename Pix> static int bd_of() { return sizeof(Pix) == 1 ? 8 : g_dropin_bd.load(); }

// hor_sad_generic / ver_sad_generic (picture-generic.c:1266-1331) replicate one reference
// row or column; build that replicated reference block on the host side of the copy and
// run the ordinary SAD kernel on it.
template <typename Pix>
static uint32_t dropin_ver_sad(const Pix *pic, const Pix *ref_row, int w, int h, uint32_t ps)
{
  return dropin_pair<Pix>(UVG_B200_SAD, bd_of<Pix>(), w, h, pic, ps, ref_row, 0);
}

template <typename Pix>
static uint32_t dropin_hor_sad(const Pix *pic, const Pix *ref, int w, int h, uint32_t ps, uint32_t rs,
                               uint32_t left, uint32_t right)
{
  if (!left && !right) return dropin_pair<Pix>(UVG_B200_SAD, bd_of<Pix>(), w, h, pic, ps, ref, rs);
  // expand the edge-replicated reference block
  Pix tmp[64 * 64];
  if (w > 64 || h > 64) { fprintf(stderr, "uvg266_b200: hor_sad block too large\n"); abort(); }
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int rx = x;
      if (left) rx = x < (int)left ? (int)left : x;                    // :1316-1320
      else rx = x >= w - (int)right ? w - (int)right - 1 : x;          // :1321-1327
      tmp[y * w + x] = ref[(size_t)y * rs + rx];
    }
  return dropin_pair<Pix>(UVG_B200_SAD, bd_of<Pix>(), w, h, pic, ps, tmp, (size_t)w);
}

template <typename Pix>
static void dropin_dual(int metric, int n, const Pix *preds, const Pix *orig, unsigned *out)
{
  const Pix *c[2] = { preds, preds + 32 * 32 };     // pred_buffer slots, strategies-picture.h:48
  dropin_distortion<Pix>(metric, bd_of<Pix>(), n, n, orig, (size_t)n, c, (size_t)n, 2, out);
}

template <typename Pix>
static void dropin_residual(const Pix *ref, const Pix *pred, int16_t *residual, int w, int h, int rs, int ps)
{
  if (ensure_init()) { fprintf(stderr, "uvg266_b200: %s\n", t_err); abort(); }
  Staging &s = tls_staging();
  const size_t blk = (size_t)w * h * sizeof(Pix);
  const size_t in_bytes = align_up(2 * blk, 16);
  const size_t out_bytes = (size_t)w * h * sizeof(int16_t);
  if (s.ensure(in_bytes + out_bytes, in_bytes + out_bytes)) { fprintf(stderr, "uvg266_b200: %s\n", t_err); abort(); }
  Pix *hp = reinterpret_cast<Pix *>(s.h_pin);
  pack_rows(hp, ref, w, h, (size_t)rs);
  pack_rows(hp + (size_t)w * h, pred, w, h, (size_t)ps);
  UVGB_MUST(cudaMemcpyAsync(s.d_buf, s.h_pin, in_bytes, cudaMemcpyHostToDevice, s.stream));
  const uvg_b200_blocks A = dense_blocks(s.d_buf, w, h), B = dense_blocks(s.d_buf + blk, w, h);
  int16_t *d_out = reinterpret_cast<int16_t *>(s.d_buf + in_bytes);
  if ((w % 4) == 0) {
    if (uvg_b200_residual_dev(bd_of<Pix>(), w, h, &A, &B, d_out, 1, s.stream)) { fprintf(stderr, "uvg266_b200: %s\n", t_err); abort(); }
  } else {
    // ISP-sized rows (w = 1, 2): treat the dense w*h block as one row of w*h pixels
    const int e = w * h;
    if (e % 4) { fprintf(stderr, "uvg266_b200: residual %dx%d unsupported\n", w, h); abort(); }
    const uvg_b200_blocks A1 = dense_blocks(s.d_buf, e, 1), B1 = dense_blocks(s.d_buf + blk, e, 1);
    if (uvg_b200_residual_dev(bd_of<Pix>(), e, 1, &A1, &B1, d_out, 1, s.stream)) { fprintf(stderr, "uvg266_b200: %s\n", t_err); abort(); }
  }
  UVGB_MUST(cudaMemcpyAsync(s.h_pin + in_bytes, d_out, out_bytes, cudaMemcpyDeviceToHost, s.stream));
  UVGB_MUST(cudaStreamSynchronize(s.stream));
  memcpy(residual, s.h_pin + in_bytes, out_bytes);
}

// transforms on one host block
static void dropin_transform(bool inverse, int bitdepth, int w, int h, int th, int tv, int sw, int sh,
                             const int16_t *in, int16_t *out)
{
  if (ensure_init()) { fprintf(stderr, "uvg266_b200: %s\n", t_err); abort(); }
  Staging &s = tls_staging();
  const size_t bytes = align_up((size_t)w * h * sizeof(int16_t), 16);
  if (s.ensure(2 * bytes, 2 * bytes)) { fprintf(stderr, "uvg266_b200: %s\n", t_err); abort(); }
  memcpy(s.h_pin, in, (size_t)w * h * sizeof(int16_t));
  UVGB_MUST(cudaMemcpyAsync(s.d_buf, s.h_pin, bytes, cudaMemcpyHostToDevice, s.stream));
  if (uvg_b200_transform_dev(inverse, bitdepth, w, h, th, tv, sw, sh, reinterpret_cast<int16_t *>(s.d_buf),
                             reinterpret_cast<int16_t *>(s.d_buf + bytes), 1, s.stream)) {
    fprintf(stderr, "uvg266_b200: %s\n", t_err); abort();
  }
  UVGB_MUST(cudaMemcpyAsync(s.h_pin + bytes, s.d_buf + bytes, bytes, cudaMemcpyDeviceToHost, s.stream));
  UVGB_MUST(cudaStreamSynchronize(s.stream));
  memcpy(out, s.h_pin + bytes, (size_t)w * h * sizeof(int16_t));
}

static void dropin_mts(bool inverse, int bitdepth, int color, int cu_type, int tr_idx, int lfnst_idx,
                       int cr_lfnst_idx, int isp_mode, int w, int h, const int16_t *in, int16_t *out,
                       int mts_type)
{
  int th, tv, sw, sh;
  derive_tr(w, h, color, cu_type, tr_idx, lfnst_idx, cr_lfnst_idx, isp_mode, mts_type, &th, &tv, &sw, &sh);
  if (th == UVG_B200_DCT2 && tv == UVG_B200_DCT2 && !lfnst_idx && !cr_lfnst_idx && w == h) { sw = 0; sh = 0; }  // :2575
  dropin_transform(inverse, bitdepth, w, h, th, tv, sw, sh, in, out);
}

template <typename Pix>
static int dropin_quantize_residual(const uvg_b200_quant_cfg *cfg, int cu_type, int tr_idx, int w, int h,
                                    int color, int use_trskip, int in_stride, int out_stride,
                                    const Pix *ref_in, const Pix *pred_in, Pix *rec_out,
                                    int16_t *coeff_out, int early_skip)
{
  if (ensure_init()) { fprintf(stderr, "uvg266_b200: %s\n", t_err); abort(); }
  Staging &s = tls_staging();
  const size_t blk = align_up((size_t)w * h * sizeof(Pix), 16);
  const size_t cbytes = align_up((size_t)w * h * sizeof(int16_t), 16);
  // device: [ref | pred | rec | coeff | has(16)]
  const size_t total = 3 * blk + cbytes + 16;
  if (s.ensure(total, total)) { fprintf(stderr, "uvg266_b200: %s\n", t_err); abort(); }
  Pix *hp = reinterpret_cast<Pix *>(s.h_pin);
  pack_rows(hp, ref_in, w, h, (size_t)in_stride);
  pack_rows(reinterpret_cast<Pix *>(s.h_pin + blk), pred_in, w, h, (size_t)in_stride);
  UVGB_MUST(cudaMemcpyAsync(s.d_buf, s.h_pin, 2 * blk, cudaMemcpyHostToDevice, s.stream));
  const uvg_b200_blocks R = dense_blocks(s.d_buf, w, h), P = dense_blocks(s.d_buf + blk, w, h),
                        O = dense_blocks(s.d_buf + 2 * blk, w, h);
  int16_t *d_coeff = reinterpret_cast<int16_t *>(s.d_buf + 3 * blk);
  uint8_t *d_has = s.d_buf + 3 * blk + cbytes;
  if (uvg_b200_quantize_residual_dev(cfg, cu_type, tr_idx, w, h, color, use_trskip, early_skip, &R, &P, &O,
                                     d_coeff, d_has, nullptr, nullptr, 0, 1, s.stream)) {
    fprintf(stderr, "uvg266_b200: %s\n", t_err); abort();
  }
  UVGB_MUST(cudaMemcpyAsync(s.h_pin + 2 * blk, s.d_buf + 2 * blk, blk + cbytes + 16, cudaMemcpyDeviceToHost, s.stream));
  UVGB_MUST(cudaStreamSynchronize(s.stream));
  const int has = s.h_pin[3 * blk + cbytes];
  memcpy(coeff_out, s.h_pin + 3 * blk, (size_t)w * h * sizeof(int16_t));
  // rec: reconstructed when coefficients survive; copy of pred otherwise unless aliased
  // (quant-generic.c:599-609)
  if ((has && !early_skip) || rec_out != pred_in) {
    const Pix *src = reinterpret_cast<const Pix *>(s.h_pin + 2 * blk);
    for (int y = 0; y < h; ++y) memcpy(rec_out + (size_t)y * out_stride, src + (size_t)y * w, (size_t)w * sizeof(Pix));
  }
  return has;
}

// uvg_quant_cbcr_residual_generic (quant-generic.c:241-442), plain chain. Returns joint_cb_cr when
// coefficients survive, else 0 (:444).
template <typename Pix>
static int dropin_quant_cbcr_residual(const uvg_b200_quant_cfg *cfg, int cu_type, int tr_idx, int w, int h,
                                      int joint_cb_cr, int jccr_sign, int in_stride, int out_stride,
                                      const Pix *u_ref, const Pix *v_ref, const Pix *u_pred, const Pix *v_pred,
                                      Pix *u_rec, Pix *v_rec, int16_t *coeff_out, int early_skip)
{
  if (ensure_init()) { fprintf(stderr, "uvg266_b200: %s\n", t_err); abort(); }
  Staging &s = tls_staging();
  const size_t blk = align_up((size_t)w * h * sizeof(Pix), 16);
  const size_t cbytes = align_up((size_t)w * h * sizeof(int16_t), 16);
  // device: [u_ref | v_ref | u_pred | v_pred | u_rec | v_rec | coeff | has(16)]
  const size_t total = 6 * blk + cbytes + 16;
  if (s.ensure(total, total)) { fprintf(stderr, "uvg266_b200: %s\n", t_err); abort(); }
  const Pix *in[4] = { u_ref, v_ref, u_pred, v_pred };
  for (int i = 0; i < 4; ++i) pack_rows(reinterpret_cast<Pix *>(s.h_pin + i * blk), in[i], w, h, (size_t)in_stride);
  UVGB_MUST(cudaMemcpyAsync(s.d_buf, s.h_pin, 4 * blk, cudaMemcpyHostToDevice, s.stream));
  uvg_b200_blocks B[6];
  for (int i = 0; i < 6; ++i) B[i] = dense_blocks(s.d_buf + i * blk, w, h);
  int16_t *d_coeff = reinterpret_cast<int16_t *>(s.d_buf + 6 * blk);
  uint8_t *d_has = s.d_buf + 6 * blk + cbytes;
  if (uvg_b200_quant_cbcr_residual_dev(cfg, cu_type, tr_idx, w, h, joint_cb_cr, jccr_sign, early_skip, &B[0], &B[1],
                                       &B[2], &B[3], &B[4], &B[5], d_coeff, d_has, 1, s.stream)) {
    fprintf(stderr, "uvg266_b200: %s\n", t_err); abort();
  }
  UVGB_MUST(cudaMemcpyAsync(s.h_pin + 4 * blk, s.d_buf + 4 * blk, 2 * blk + cbytes + 16, cudaMemcpyDeviceToHost, s.stream));
  UVGB_MUST(cudaStreamSynchronize(s.stream));
  const int has = s.h_pin[6 * blk + cbytes];
  memcpy(coeff_out, s.h_pin + 6 * blk, (size_t)w * h * sizeof(int16_t));
  const Pix *su = reinterpret_cast<const Pix *>(s.h_pin + 4 * blk), *sv = reinterpret_cast<const Pix *>(s.h_pin + 5 * blk);
  for (int y = 0; y < h; ++y) {                     // both planes are always written (:415-441)
    memcpy(u_rec + (size_t)y * out_stride, su + (size_t)y * w, (size_t)w * sizeof(Pix));
    memcpy(v_rec + (size_t)y * out_stride, sv + (size_t)y * w, (size_t)w * sizeof(Pix));
  }
  return has ? joint_cb_cr : 0;
}

}  // namespace uvgb

using namespace uvgb;

// ---------------------------------------------------------------------------------------
// library / device
// ---------------------------------------------------------------------------------------
extern "C" int uvg_b200_init(int device) { return do_init(device); }

extern "C" int uvg_b200_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" const char *uvg_b200_last_error(void) { return t_err; }
extern "C" const char *uvg_b200_version(void) { return "uvg266_b200 0.1 (sm_100a)"; }
extern "C" int uvg_b200_sm_count(void) { return ensure_init() ? 0 : g_sms; }
extern "C" uint64_t uvg_b200_launch_count(void) { return g_launches.load(); }
extern "C" void uvg_b200_set_dropin_bitdepth(int bitdepth) { g_dropin_bd.store(bitdepth); }

extern "C" void *uvg_b200_host_alloc(size_t bytes)
{
  if (ensure_init()) return nullptr;
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) { set_error("cudaMallocHost(%zu) failed", bytes); cudaGetLastError(); return nullptr; }
  return p;
}

extern "C" void uvg_b200_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---------------------------------------------------------------------------------------
// batched host entry points
// ---------------------------------------------------------------------------------------
extern "C" int uvg_b200_distortion_host(int metric, int bitdepth, int w, int h, const uvg_b200_blocks *a,
                                        size_t a_bytes, const uvg_b200_blocks *b, size_t b_bytes,
                                        uint32_t *out, size_t n_blocks, int k)
{
  if (ensure_init()) return 1;
  if (n_blocks == 0) return 0;
  Staging &s = tls_staging();
  const size_t ao = 0, bo = align_up(a_bytes, 256), oo = bo + align_up(b_bytes, 256);
  const size_t out_bytes = n_blocks * (size_t)k * sizeof(uint32_t);
  if (s.ensure(0, oo + out_bytes)) return 1;
  UVGB_TRY(cudaMemcpyAsync(s.d_buf + ao, a->base, a_bytes, cudaMemcpyHostToDevice, s.stream));
  UVGB_TRY(cudaMemcpyAsync(s.d_buf + bo, b->base, b_bytes, cudaMemcpyHostToDevice, s.stream));
  uvg_b200_blocks A = *a, B = *b;
  A.base = s.d_buf + ao; B.base = s.d_buf + bo;
  uint32_t *d_out = reinterpret_cast<uint32_t *>(s.d_buf + oo);
  if (uvg_b200_distortion_dev(metric, bitdepth, w, h, &A, &B, d_out, n_blocks, k, s.stream)) return 1;
  UVGB_TRY(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, s.stream));
  UVGB_TRY(cudaStreamSynchronize(s.stream));
  return 0;
}

extern "C" int uvg_b200_rough_cost_host(int bitdepth, int n, const uvg_b200_blocks *orig, size_t orig_bytes,
                                        const uvg_b200_blocks *preds, size_t preds_bytes, uint32_t *out,
                                        size_t n_blocks, int k)
{
  if (ensure_init()) return 1;
  if (n_blocks == 0) return 0;
  Staging &s = tls_staging();
  const size_t bo = align_up(orig_bytes, 256), oo = bo + align_up(preds_bytes, 256);
  const size_t out_bytes = n_blocks * (size_t)k * sizeof(uint32_t);
  if (s.ensure(0, oo + out_bytes)) return 1;
  UVGB_TRY(cudaMemcpyAsync(s.d_buf, orig->base, orig_bytes, cudaMemcpyHostToDevice, s.stream));
  UVGB_TRY(cudaMemcpyAsync(s.d_buf + bo, preds->base, preds_bytes, cudaMemcpyHostToDevice, s.stream));
  uvg_b200_blocks A = *orig, B = *preds;
  A.base = s.d_buf; B.base = s.d_buf + bo;
  uint32_t *d_out = reinterpret_cast<uint32_t *>(s.d_buf + oo);
  if (uvg_b200_rough_cost_dev(bitdepth, n, &A, &B, d_out, n_blocks, k, s.stream)) return 1;
  UVGB_TRY(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, s.stream));
  UVGB_TRY(cudaStreamSynchronize(s.stream));
  return 0;
}

extern "C" int uvg_b200_intra_predict_host(int bitdepth, int n, const void *refs, const int8_t *modes, int k,
                                           void *preds, size_t n_blocks)
{
  if (ensure_init()) return 1;
  if (n_blocks == 0) return 0;
  if (n < 4 || n > 32 || k < 1 || k > 67) { set_error("intra_predict_host: bad n=%d or k=%d", n, k); return 1; }
  Staging &s = tls_staging();
  const size_t pix = bitdepth == 8 ? 1 : 2;
  const size_t ref_bytes = n_blocks * (size_t)(2 * (2 * n + 4)) * pix, rb = align_up(ref_bytes, 256);
  const size_t out_bytes = n_blocks * (size_t)k * n * n * pix;
  if (s.ensure(0, rb + out_bytes)) return 1;
  UVGB_TRY(cudaMemcpyAsync(s.d_buf, refs, ref_bytes, cudaMemcpyHostToDevice, s.stream));
  if (uvg_b200_intra_predict_dev(bitdepth, n, s.d_buf, modes, k, s.d_buf + rb, n_blocks, s.stream)) return 1;
  UVGB_TRY(cudaMemcpyAsync(preds, s.d_buf + rb, out_bytes, cudaMemcpyDeviceToHost, s.stream));
  UVGB_TRY(cudaStreamSynchronize(s.stream));
  return 0;
}

extern "C" int uvg_b200_intra_rough_host(int bitdepth, int n, const uvg_b200_blocks *orig, size_t orig_bytes,
                                         const void *refs, const int8_t *modes, int k, uint32_t *costs,
                                         size_t n_blocks)
{
  if (ensure_init()) return 1;
  if (n_blocks == 0) return 0;
  if (n < 4 || n > 32 || k < 1 || k > 67) { set_error("intra_rough_host: bad n=%d or k=%d", n, k); return 1; }
  Staging &s = tls_staging();
  const size_t pix = bitdepth == 8 ? 1 : 2;
  const size_t ref_bytes = n_blocks * (size_t)(2 * (2 * n + 4)) * pix;
  const size_t bo = align_up(orig_bytes, 256), ro = bo + align_up(ref_bytes, 256);
  const size_t out_bytes = n_blocks * (size_t)k * sizeof(uint32_t);
  if (s.ensure(0, ro + out_bytes)) return 1;
  UVGB_TRY(cudaMemcpyAsync(s.d_buf, orig->base, orig_bytes, cudaMemcpyHostToDevice, s.stream));
  UVGB_TRY(cudaMemcpyAsync(s.d_buf + bo, refs, ref_bytes, cudaMemcpyHostToDevice, s.stream));
  uvg_b200_blocks A = *orig;
  A.base = s.d_buf;
  uint32_t *d_out = reinterpret_cast<uint32_t *>(s.d_buf + ro);
  if (uvg_b200_intra_rough_dev(bitdepth, n, &A, s.d_buf + bo, modes, k, d_out, n_blocks, s.stream)) return 1;
  UVGB_TRY(cudaMemcpyAsync(costs, d_out, out_bytes, cudaMemcpyDeviceToHost, s.stream));
  UVGB_TRY(cudaStreamSynchronize(s.stream));
  return 0;
}

extern "C" int uvg_b200_transform_host(int inverse, int bitdepth, int w, int h, int type_hor, int type_ver,
                                       int skip_w, int skip_h, const int16_t *in, int16_t *out, size_t n_tus)
{
  if (ensure_init()) return 1;
  if (n_tus == 0) return 0;
  Staging &s = tls_staging();
  const size_t bytes = align_up(n_tus * (size_t)w * h * sizeof(int16_t), 256);
  if (s.ensure(0, 2 * bytes)) return 1;
  UVGB_TRY(cudaMemcpyAsync(s.d_buf, in, n_tus * (size_t)w * h * sizeof(int16_t), cudaMemcpyHostToDevice, s.stream));
  if (uvg_b200_transform_dev(inverse, bitdepth, w, h, type_hor, type_ver, skip_w, skip_h,
                             reinterpret_cast<int16_t *>(s.d_buf), reinterpret_cast<int16_t *>(s.d_buf + bytes),
                             n_tus, s.stream)) return 1;
  UVGB_TRY(cudaMemcpyAsync(out, s.d_buf + bytes, n_tus * (size_t)w * h * sizeof(int16_t), cudaMemcpyDeviceToHost, s.stream));
  UVGB_TRY(cudaStreamSynchronize(s.stream));
  return 0;
}

extern "C" int uvg_b200_quantize_residual_host(const uvg_b200_quant_cfg *cfg, int cu_type, int tr_idx, int w,
                                               int h, int color, int use_trskip, int early_skip,
                                               const uvg_b200_blocks *ref, size_t ref_bytes,
                                               const uvg_b200_blocks *pred, size_t pred_bytes,
                                               const uvg_b200_blocks *rec, size_t rec_bytes,
                                               int16_t *coeff_out, uint8_t *has_coeffs, uint32_t *ssd,
                                               uint32_t *cost, uint64_t cost_weights, size_t n_tus)
{
  if (ensure_init()) return 1;
  if (n_tus == 0) return 0;
  Staging &s = tls_staging();
  const bool alias = rec->base == pred->base;
  const size_t po = align_up(ref_bytes, 256);
  const size_t ro = alias ? po : po + align_up(pred_bytes, 256);
  const size_t co = ro + align_up(alias ? pred_bytes : rec_bytes, 256);
  const size_t cbytes = n_tus * (size_t)w * h * sizeof(int16_t);
  const size_t ho = co + align_up(cbytes, 256);
  const size_t so = ho + align_up(n_tus, 256);
  const size_t to = so + align_up(n_tus * 4, 256);
  const size_t total = to + align_up(n_tus * 4, 256);
  if (s.ensure(0, total)) return 1;
  UVGB_TRY(cudaMemcpyAsync(s.d_buf, ref->base, ref_bytes, cudaMemcpyHostToDevice, s.stream));
  UVGB_TRY(cudaMemcpyAsync(s.d_buf + po, pred->base, pred_bytes, cudaMemcpyHostToDevice, s.stream));
  uvg_b200_blocks R = *ref, P = *pred, O = *rec;
  R.base = s.d_buf; P.base = s.d_buf + po; O.base = s.d_buf + ro;
  if (uvg_b200_quantize_residual_dev(cfg, cu_type, tr_idx, w, h, color, use_trskip, early_skip, &R, &P, &O,
                                     reinterpret_cast<int16_t *>(s.d_buf + co), s.d_buf + ho,
                                     ssd ? reinterpret_cast<uint32_t *>(s.d_buf + so) : nullptr,
                                     cost ? reinterpret_cast<uint32_t *>(s.d_buf + to) : nullptr, cost_weights,
                                     n_tus, s.stream)) return 1;
  UVGB_TRY(cudaMemcpyAsync(const_cast<void *>(rec->base), s.d_buf + ro, alias ? pred_bytes : rec_bytes,
                           cudaMemcpyDeviceToHost, s.stream));
  UVGB_TRY(cudaMemcpyAsync(coeff_out, s.d_buf + co, cbytes, cudaMemcpyDeviceToHost, s.stream));
  UVGB_TRY(cudaMemcpyAsync(has_coeffs, s.d_buf + ho, n_tus, cudaMemcpyDeviceToHost, s.stream));
  if (ssd) UVGB_TRY(cudaMemcpyAsync(ssd, s.d_buf + so, n_tus * 4, cudaMemcpyDeviceToHost, s.stream));
  if (cost) UVGB_TRY(cudaMemcpyAsync(cost, s.d_buf + to, n_tus * 4, cudaMemcpyDeviceToHost, s.stream));
  UVGB_TRY(cudaStreamSynchronize(s.stream));
  return 0;
}

// ---------------------------------------------------------------------------------------
// drop-in leaf functions
// ---------------------------------------------------------------------------------------
#define UVGB_DEF_PICTURE(SUF, PIX)                                                                         \
  extern "C" unsigned uvg_b200_reg_sad_##SUF(const PIX *d1, const PIX *d2, int w, int h, unsigned s1,      \
                                             unsigned s2)                                                  \
  { return dropin_pair<PIX>(UVG_B200_SAD, bd_of<PIX>(), w, h, d1, s1, d2, s2); }                          \
  extern "C" unsigned uvg_b200_sad_4x4_##SUF(const PIX *a, const PIX *b) { return dropin_pair<PIX>(UVG_B200_SAD_NXN, bd_of<PIX>(), 4, 4, a, 4, b, 4); }       \
  extern "C" unsigned uvg_b200_sad_8x8_##SUF(const PIX *a, const PIX *b) { return dropin_pair<PIX>(UVG_B200_SAD_NXN, bd_of<PIX>(), 8, 8, a, 8, b, 8); }       \
  extern "C" unsigned uvg_b200_sad_16x16_##SUF(const PIX *a, const PIX *b) { return dropin_pair<PIX>(UVG_B200_SAD_NXN, bd_of<PIX>(), 16, 16, a, 16, b, 16); } \
  extern "C" unsigned uvg_b200_sad_32x32_##SUF(const PIX *a, const PIX *b) { return dropin_pair<PIX>(UVG_B200_SAD_NXN, bd_of<PIX>(), 32, 32, a, 32, b, 32); } \
  extern "C" unsigned uvg_b200_sad_64x64_##SUF(const PIX *a, const PIX *b) { return dropin_pair<PIX>(UVG_B200_SAD_NXN, bd_of<PIX>(), 64, 64, a, 64, b, 64); } \
  extern "C" unsigned uvg_b200_satd_4x4_##SUF(const PIX *a, const PIX *b) { return dropin_pair<PIX>(UVG_B200_SATD_4X4, bd_of<PIX>(), 4, 4, a, 4, b, 4); }     \
  extern "C" unsigned uvg_b200_satd_8x8_##SUF(const PIX *a, const PIX *b) { return dropin_pair<PIX>(UVG_B200_SATD, bd_of<PIX>(), 8, 8, a, 8, b, 8); }         \
  extern "C" unsigned uvg_b200_satd_16x16_##SUF(const PIX *a, const PIX *b) { return dropin_pair<PIX>(UVG_B200_SATD, bd_of<PIX>(), 16, 16, a, 16, b, 16); }   \
  extern "C" unsigned uvg_b200_satd_32x32_##SUF(const PIX *a, const PIX *b) { return dropin_pair<PIX>(UVG_B200_SATD, bd_of<PIX>(), 32, 32, a, 32, b, 32); }   \
  extern "C" unsigned uvg_b200_satd_64x64_##SUF(const PIX *a, const PIX *b) { return dropin_pair<PIX>(UVG_B200_SATD, bd_of<PIX>(), 64, 64, a, 64, b, 64); }   \
  extern "C" void uvg_b200_sad_4x4_dual_##SUF(const PIX *p, const PIX *o, unsigned n, unsigned *out) { (void)n; dropin_dual<PIX>(UVG_B200_SAD_NXN, 4, p, o, out); }     \
  extern "C" void uvg_b200_sad_8x8_dual_##SUF(const PIX *p, const PIX *o, unsigned n, unsigned *out) { (void)n; dropin_dual<PIX>(UVG_B200_SAD_NXN, 8, p, o, out); }     \
  extern "C" void uvg_b200_sad_16x16_dual_##SUF(const PIX *p, const PIX *o, unsigned n, unsigned *out) { (void)n; dropin_dual<PIX>(UVG_B200_SAD_NXN, 16, p, o, out); }  \
  extern "C" void uvg_b200_sad_32x32_dual_##SUF(const PIX *p, const PIX *o, unsigned n, unsigned *out) { (void)n; dropin_dual<PIX>(UVG_B200_SAD_NXN, 32, p, o, out); }  \
  extern "C" void uvg_b200_sad_64x64_dual_##SUF(const PIX *p, const PIX *o, unsigned n, unsigned *out) { (void)n; dropin_dual<PIX>(UVG_B200_SAD_NXN, 64, p, o, out); }  \
  extern "C" void uvg_b200_satd_64x64_dual_##SUF(const PIX *p, const PIX *o, unsigned n, unsigned *out) { (void)n; dropin_dual<PIX>(UVG_B200_SATD, 64, p, o, out); }    \
  extern "C" void uvg_b200_satd_4x4_dual_##SUF(const PIX *p, const PIX *o, unsigned n, unsigned *out) { (void)n; dropin_dual<PIX>(UVG_B200_SATD_4X4, 4, p, o, out); }   \
  extern "C" void uvg_b200_satd_8x8_dual_##SUF(const PIX *p, const PIX *o, unsigned n, unsigned *out) { (void)n; dropin_dual<PIX>(UVG_B200_SATD, 8, p, o, out); }       \
  extern "C" void uvg_b200_satd_16x16_dual_##SUF(const PIX *p, const PIX *o, unsigned n, unsigned *out) { (void)n; dropin_dual<PIX>(UVG_B200_SATD, 16, p, o, out); }    \
  extern "C" void uvg_b200_satd_32x32_dual_##SUF(const PIX *p, const PIX *o, unsigned n, unsigned *out) { (void)n; dropin_dual<PIX>(UVG_B200_SATD, 32, p, o, out); }    \
  extern "C" unsigned uvg_b200_satd_any_size_##SUF(int w, int h, const PIX *b1, int s1, const PIX *b2, int s2)            \
  { return dropin_pair<PIX>(UVG_B200_SATD, bd_of<PIX>(), w, h, b1, (size_t)s1, b2, (size_t)s2); }                       \
  extern "C" unsigned uvg_b200_satd_any_size_vtm_##SUF(int w, int h, const PIX *b1, int s1, const PIX *b2, int s2)        \
  { return dropin_pair<PIX>(UVG_B200_SATD_VTM, bd_of<PIX>(), w, h, b1, (size_t)s1, b2, (size_t)s2); }                   \
  extern "C" void uvg_b200_satd_any_size_quad_##SUF(int w, int h, const PIX **preds, int stride, const PIX *orig,         \
                                                    int orig_stride, unsigned num_modes, unsigned *costs_out,            \
                                                    int8_t *valid)                                                       \
  { (void)num_modes; (void)valid;                                                                                        \
    dropin_distortion<PIX>(UVG_B200_SATD_QUAD, bd_of<PIX>(), w, h, orig, (size_t)orig_stride, preds, (size_t)stride, 4, costs_out); } \
  extern "C" unsigned uvg_b200_pixels_calc_ssd_##SUF(const PIX *ref, const PIX *rec, int rs, int cs, int w, int h)        \
  { return dropin_pair<PIX>(UVG_B200_SSD, bd_of<PIX>(), w, h, ref, (size_t)rs, rec, (size_t)cs); }                      \
  extern "C" uint32_t uvg_b200_ver_sad_##SUF(const PIX *pic, const PIX *ref_row, int32_t w, int32_t h, uint32_t ps)       \
  { return dropin_ver_sad<PIX>(pic, ref_row, w, h, ps); }                                                                \
  extern "C" uint32_t uvg_b200_hor_sad_##SUF(const PIX *pic, const PIX *ref, int32_t w, int32_t h, uint32_t ps,           \
                                             uint32_t rs, uint32_t left, uint32_t right)                                 \
  { return dropin_hor_sad<PIX>(pic, ref, w, h, ps, rs, left, right); }                                                   \
  extern "C" void uvg_b200_generate_residual_##SUF(const PIX *ref, const PIX *pred, int16_t *residual, int w, int h,      \
                                                   int rs, int ps)                                                       \
  { dropin_residual<PIX>(ref, pred, residual, w, h, rs, ps); }

UVGB_DEF_PICTURE(8, uint8_t)
UVGB_DEF_PICTURE(16, uint16_t)

// get_optimized_sad (strategies-picture.h:128): width-specialised reg_sad variants. The
// reference's AVX2 strategy returns functions for widths 4/8/12/16/24/32/64
// (picture-avx2.c:1497-1517); every one must equal reg_sad for its width.
#define UVGB_OPT_SAD(SUF, PIX, W)                                                                       \
  static uint32_t opt_sad_##W##_##SUF(const PIX *a, const PIX *b, int32_t h, uint32_t s1, uint32_t s2)   \
  { return dropin_pair<PIX>(UVG_B200_SAD, bd_of<PIX>(), W, h, a, s1, b, s2); }
#define UVGB_OPT_SAD_SET(SUF, PIX)                                                                      \
  UVGB_OPT_SAD(SUF, PIX, 4) UVGB_OPT_SAD(SUF, PIX, 8) UVGB_OPT_SAD(SUF, PIX, 12) UVGB_OPT_SAD(SUF, PIX, 16) \
  UVGB_OPT_SAD(SUF, PIX, 24) UVGB_OPT_SAD(SUF, PIX, 32) UVGB_OPT_SAD(SUF, PIX, 64)                      \
  extern "C" uvg_b200_optimized_sad_##SUF uvg_b200_get_optimized_sad_##SUF(int32_t width)               \
  {                                                                                                     \
    switch (width) {                                                                                    \
      case 4: return opt_sad_4_##SUF; case 8: return opt_sad_8_##SUF; case 12: return opt_sad_12_##SUF; \
      case 16: return opt_sad_16_##SUF; case 24: return opt_sad_24_##SUF; case 32: return opt_sad_32_##SUF; \
      case 64: return opt_sad_64_##SUF; default: return nullptr;                                        \
    }                                                                                                   \
  }
UVGB_OPT_SAD_SET(8, uint8_t)
UVGB_OPT_SAD_SET(16, uint16_t)

#define UVGB_DEF_DCT(N)                                                                                   \
  extern "C" void uvg_b200_dct_##N##x##N(int8_t bd, const int16_t *in, int16_t *out)                       \
  { dropin_transform(false, bd, N, N, UVG_B200_DCT2, UVG_B200_DCT2, 0, 0, in, out); }                     \
  extern "C" void uvg_b200_idct_##N##x##N(int8_t bd, const int16_t *in, int16_t *out)                      \
  { dropin_transform(true, bd, N, N, UVG_B200_DCT2, UVG_B200_DCT2, 0, 0, in, out); }
UVGB_DEF_DCT(4) UVGB_DEF_DCT(8) UVGB_DEF_DCT(16) UVGB_DEF_DCT(32)

// fast_forward/inverse_dst_4x4 (dct-generic.c:752-770): 4x4 DST-7 in both directions with
// the dct_4x4 shifts.
extern "C" void uvg_b200_fast_forward_dst_4x4(int8_t bd, const int16_t *in, int16_t *out)
{ dropin_transform(false, bd, 4, 4, UVG_B200_DST7, UVG_B200_DST7, 0, 0, in, out); }
extern "C" void uvg_b200_fast_inverse_dst_4x4(int8_t bd, const int16_t *in, int16_t *out)
{ dropin_transform(true, bd, 4, 4, UVG_B200_DST7, UVG_B200_DST7, 0, 0, in, out); }

extern "C" void uvg_b200_mts_dct(int8_t bd, int color, int cu_type, int tr_idx, int lfnst_idx, int cr_lfnst_idx,
                                 int isp_mode, int8_t w, int8_t h, const int16_t *in, int16_t *out, int8_t mts_type)
{ dropin_mts(false, bd, color, cu_type, tr_idx, lfnst_idx, cr_lfnst_idx, isp_mode, w, h, in, out, mts_type); }
extern "C" void uvg_b200_mts_idct(int8_t bd, int color, int cu_type, int tr_idx, int lfnst_idx, int cr_lfnst_idx,
                                  int isp_mode, int8_t w, int8_t h, const int16_t *in, int16_t *out, int8_t mts_type)
{ dropin_mts(true, bd, color, cu_type, tr_idx, lfnst_idx, cr_lfnst_idx, isp_mode, w, h, in, out, mts_type); }

static void dropin_quant_common(bool deq, const uvg_b200_quant_cfg *cfg, const int16_t *in, int16_t *out, int w,
                                int h, int color, int ts, int lfnst_idx)
{
  if (ensure_init()) { fprintf(stderr, "uvg266_b200: %s\n", uvg_b200_last_error()); abort(); }
  Staging &s = tls_staging();
  const size_t bytes = align_up((size_t)w * h * sizeof(int16_t), 16);
  if (s.ensure(2 * bytes, 2 * bytes)) { fprintf(stderr, "uvg266_b200: %s\n", uvg_b200_last_error()); abort(); }
  memcpy(s.h_pin, in, (size_t)w * h * sizeof(int16_t));
  UVGB_MUST(cudaMemcpyAsync(s.d_buf, s.h_pin, bytes, cudaMemcpyHostToDevice, s.stream));
  int rc = deq ? uvg_b200_dequant_dev(cfg, reinterpret_cast<int16_t *>(s.d_buf), reinterpret_cast<int16_t *>(s.d_buf + bytes),
                                      w, h, color, ts, 1, s.stream)
               : uvg_b200_quant_dev(cfg, reinterpret_cast<int16_t *>(s.d_buf), reinterpret_cast<int16_t *>(s.d_buf + bytes),
                                    w, h, color, ts, lfnst_idx, 1, s.stream);
  if (rc) { fprintf(stderr, "uvg266_b200: %s\n", uvg_b200_last_error()); abort(); }
  UVGB_MUST(cudaMemcpyAsync(s.h_pin + bytes, s.d_buf + bytes, bytes, cudaMemcpyDeviceToHost, s.stream));
  UVGB_MUST(cudaStreamSynchronize(s.stream));
  memcpy(out, s.h_pin + bytes, (size_t)w * h * sizeof(int16_t));
}

extern "C" void uvg_b200_quant(const uvg_b200_quant_cfg *cfg, const int16_t *coef, int16_t *q_coef, int32_t w,
                               int32_t h, int color, int8_t scan_idx, int8_t block_type, int8_t transform_skip,
                               uint8_t lfnst_idx)
{ (void)scan_idx; (void)block_type; dropin_quant_common(false, cfg, coef, q_coef, w, h, color, transform_skip, lfnst_idx); }

extern "C" void uvg_b200_dequant(const uvg_b200_quant_cfg *cfg, const int16_t *q_coef, int16_t *coef, int32_t w,
                                 int32_t h, int color, int8_t block_type, int8_t transform_skip)
{ (void)block_type; dropin_quant_common(true, cfg, q_coef, coef, w, h, color, transform_skip, 0); }

extern "C" int uvg_b200_quantize_residual_8(const uvg_b200_quant_cfg *cfg, int cu_type, int tr_idx, int w, int h,
                                            int color, int use_trskip, int in_stride, int out_stride,
                                            const uint8_t *ref_in, const uint8_t *pred_in, uint8_t *rec_out,
                                            int16_t *coeff_out, int early_skip)
{ return dropin_quantize_residual<uint8_t>(cfg, cu_type, tr_idx, w, h, color, use_trskip, in_stride, out_stride, ref_in, pred_in, rec_out, coeff_out, early_skip); }

extern "C" int uvg_b200_quantize_residual_16(const uvg_b200_quant_cfg *cfg, int cu_type, int tr_idx, int w, int h,
                                             int color, int use_trskip, int in_stride, int out_stride,
                                             const uint16_t *ref_in, const uint16_t *pred_in, uint16_t *rec_out,
                                             int16_t *coeff_out, int early_skip)
{ return dropin_quantize_residual<uint16_t>(cfg, cu_type, tr_idx, w, h, color, use_trskip, in_stride, out_stride, ref_in, pred_in, rec_out, coeff_out, early_skip); }

extern "C" int uvg_b200_quant_cbcr_residual_8(const uvg_b200_quant_cfg *cfg, int cu_type, int tr_idx, int w, int h,
                                              int joint_cb_cr, int jccr_sign, int in_stride, int out_stride,
                                              const uint8_t *u_ref, const uint8_t *v_ref, const uint8_t *u_pred,
                                              const uint8_t *v_pred, uint8_t *u_rec, uint8_t *v_rec,
                                              int16_t *coeff_out, int early_skip)
{ return dropin_quant_cbcr_residual<uint8_t>(cfg, cu_type, tr_idx, w, h, joint_cb_cr, jccr_sign, in_stride, out_stride, u_ref, v_ref, u_pred, v_pred, u_rec, v_rec, coeff_out, early_skip); }

extern "C" int uvg_b200_quant_cbcr_residual_16(const uvg_b200_quant_cfg *cfg, int cu_type, int tr_idx, int w, int h,
                                               int joint_cb_cr, int jccr_sign, int in_stride, int out_stride,
                                               const uint16_t *u_ref, const uint16_t *v_ref, const uint16_t *u_pred,
                                               const uint16_t *v_pred, uint16_t *u_rec, uint16_t *v_rec,
                                               int16_t *coeff_out, int early_skip)
{ return dropin_quant_cbcr_residual<uint16_t>(cfg, cu_type, tr_idx, w, h, joint_cb_cr, jccr_sign, in_stride, out_stride, u_ref, v_ref, u_pred, v_pred, u_rec, v_rec, coeff_out, early_skip); }

static uint32_t dropin_coeff_sum(const int16_t *coeffs, size_t len, int mode, int w, int h, uint64_t weights)
{
  if (ensure_init()) { fprintf(stderr, "uvg266_b200: %s\n", uvg_b200_last_error()); abort(); }
  Staging &s = tls_staging();
  const size_t bytes = align_up(len * sizeof(int16_t), 16);
  if (s.ensure(bytes + 16, bytes + 16)) { fprintf(stderr, "uvg266_b200: %s\n", uvg_b200_last_error()); abort(); }
  memcpy(s.h_pin, coeffs, len * sizeof(int16_t));
  UVGB_MUST(cudaMemcpyAsync(s.d_buf, s.h_pin, bytes, cudaMemcpyHostToDevice, s.stream));
  uint32_t *d_out = reinterpret_cast<uint32_t *>(s.d_buf + bytes);
  int rc = mode == 0 ? uvg_b200_coeff_abs_sum_dev(reinterpret_cast<int16_t *>(s.d_buf), len, d_out, 1, s.stream)
                     : uvg_b200_fast_coeff_cost_dev(reinterpret_cast<int16_t *>(s.d_buf), w, h, weights, d_out, 1, s.stream);
  if (rc) { fprintf(stderr, "uvg266_b200: %s\n", uvg_b200_last_error()); abort(); }
  UVGB_MUST(cudaMemcpyAsync(s.h_pin + bytes, d_out, 4, cudaMemcpyDeviceToHost, s.stream));
  UVGB_MUST(cudaStreamSynchronize(s.stream));
  uint32_t r;
  memcpy(&r, s.h_pin + bytes, 4);
  return r;
}

extern "C" uint32_t uvg_b200_coeff_abs_sum(const int16_t *coeffs, size_t length)
{ return dropin_coeff_sum(coeffs, length, 0, 0, 0, 0); }
extern "C" uint32_t uvg_b200_fast_coeff_cost(const int16_t *coeff, int32_t w, int32_t h, uint64_t weights)
{ return dropin_coeff_sum(coeff, (size_t)w * h, 1, w, h, weights); }
