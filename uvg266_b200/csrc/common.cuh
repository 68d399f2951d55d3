// common.cuh - shared device/host helpers for libuvg266_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <atomic>
#include "../../include/uvg266_b200.h"

namespace uvgb {

// ---- error handling ---------------------------------------------------------------
void set_error(const char *fmt, ...);
[[noreturn]] void fatal(const char *what, cudaError_t e, const char *file, int line);
extern std::atomic<uint64_t> g_launches;

#define UVGB_TRY(expr)                                                          \
  do {                                                                          \
    cudaError_t e__ = (expr);                                                   \
    if (e__ != cudaSuccess) {                                                   \
      ::uvgb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                    \
      return 1;                                                                 \
    }                                                                           \
  } while (0)

// Leaf drop-ins have no error channel (strategies-*.h): abort loudly, never fall back.
#define UVGB_MUST(expr)                                                 \
  do {                                                                  \
    cudaError_t e__ = (expr);                                           \
    if (e__ != cudaSuccess) ::uvgb::fatal(#expr, e__, __FILE__, __LINE__); \
  } while (0)

#define UVGB_LAUNCH_CHECK()                                                    \
  do {                                                                         \
    ::uvgb::g_launches.fetch_add(1, std::memory_order_relaxed);                \
    cudaError_t e__ = cudaGetLastError();                                      \
    if (e__ != cudaSuccess) {                                                  \
      ::uvgb::set_error("kernel launch failed: %s (%s:%d)",                    \
                        cudaGetErrorString(e__), __FILE__, __LINE__);          \
      return 1;                                                                \
    }                                                                          \
  } while (0)

int ensure_init();      // 0 ok; sets device, tables
int dropin_bitdepth(size_t pixel_size);   // bit depth the leaf drop-ins assume for this pixel type (api.cu)
int sm_count();

// ---- block addressing ---------------------------------------------------------------
struct Blocks {          // device-side copy of uvg_b200_blocks
  const void *base;
  uint32_t stride;
  uint32_t bpr;
  uint64_t col_pitch;
  uint64_t row_pitch;
  __host__ __device__ __forceinline__ uint64_t offset(uint64_t i) const {
    if (bpr == 1) return i * row_pitch;
    return (i / bpr) * row_pitch + (i % bpr) * col_pitch;
  }
};

static inline Blocks to_blocks(const uvg_b200_blocks *b)
{
  Blocks r;
  r.base = b->base; r.stride = b->stride; r.bpr = b->blocks_per_row ? b->blocks_per_row : 1;
  r.col_pitch = b->col_pitch; r.row_pitch = b->row_pitch;
  return r;
}

// Bytes an operand spans for n blocks of w*h pixels of `pix` bytes (for host staging).
static inline size_t blocks_span_bytes(const uvg_b200_blocks *b, size_t n, int w, int h, int pix)
{
  if (n == 0) return 0;
  const Blocks bl = to_blocks(b);
  uint64_t last;
  if (bl.bpr == 1) {
    last = (n - 1) * bl.row_pitch;
  } else {
    const uint64_t rfull = n / bl.bpr, rem = n % bl.bpr;
    last = 0;
    if (rfull > 0) last = (rfull - 1) * bl.row_pitch + (uint64_t)(bl.bpr - 1) * bl.col_pitch;
    if (rem > 0) {
      const uint64_t c = rfull * bl.row_pitch + (rem - 1) * bl.col_pitch;
      if (c > last) last = c;
    }
  }
  return (size_t)((last + (uint64_t)(h - 1) * bl.stride + (uint64_t)w) * (uint64_t)pix);
}

// ---- small device helpers ---------------------------------------------------------
__device__ __forceinline__ int warp_sum(int v)
{
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

// sum over the `G` lanes (power of two, lanes contiguous and G-aligned) of a warp
template <int G>
__device__ __forceinline__ unsigned group_sum(unsigned v)
{
#pragma unroll
  for (int d = G / 2; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// four values -> four bytes, each saturated to 0..255 (two I2IP: clip and pack in one)
__device__ __forceinline__ uint32_t pack4_sat_u8(int v0, int v1, int v2, int v3)
{
  uint32_t d;
  asm("{ .reg .b32 t;\n\t"
      "cvt.pack.sat.u8.s32.b32 t, %4, %3, 0;\n\t"
      "cvt.pack.sat.u8.s32.b32 %0, %2, %1, t; }"
      : "=r"(d) : "r"(v0), "r"(v1), "r"(v2), "r"(v3));
  return d;
}

// two values -> two int16 halves (lo in bits 0-15), each saturated to -32768..32767: clip3(-32768, 32767, .) of both and
// the pair packing of a byte-plane split in ONE instruction
__device__ __forceinline__ uint32_t pack2_sat_s16(int lo, int hi)
{
  uint32_t d;
  asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(d) : "r"(hi), "r"(lo));
  return d;
}

__device__ __forceinline__ int clip3(int lo, int hi, int v) { return min(max(v, lo), hi); }

static inline int ilog2_host(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// Transform matrices, generated at init from the VVC base constants (transform.cu).
// Layout: g_tr[type][log2n-1] -> n*n int8 row-major T[k*n+i]; all |c| <= 90.
const int8_t *tr_matrix_dev(int type, int n);
const int8_t *tr_matrix_host(int type, int n);

// ---- per-thread staging for the host entry points (api.cu) --------------------------
struct Staging {
  cudaStream_t stream = nullptr;
  uint8_t *h_pin = nullptr; size_t h_cap = 0;   // pinned host
  uint8_t *d_buf = nullptr; size_t d_cap = 0;   // device
  int ensure(size_t host_bytes, size_t dev_bytes);
  ~Staging();     // releases the thread's stream and buffers when an encoder worker thread exits
};
Staging &tls_staging();

}  // namespace uvgb
