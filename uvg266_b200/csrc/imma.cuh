// imma.cuh - the mma.sync INT8 building blocks shared by the fused quantize_residual kernels (fused_imma.cu) and the
// standalone transform kernels (transform_imma.cu): the instruction wrapper, the byte-plane split of 16-bit values
// and the contraction-slot permutation that lets a C fragment be the next product's A or B fragment.
#pragma once
#include "common.cuh"

namespace uvgb {

// D(16x8,s32) += A(16x16 bytes) * B(16x8 bytes); AU/BU: operand is unsigned
template <bool AU, bool BU>
__device__ __forceinline__ void imma(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0)
{
  if (AU && !BU)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
  else if (!AU && BU)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
}

// bytes (lo plane, hi plane) of four 16-bit values held in int registers
__device__ __forceinline__ void split4(int v0, int v1, int v2, int v3, uint32_t &lo, uint32_t &hi)
{
  const uint32_t p01 = __byte_perm((uint32_t)v0, (uint32_t)v1, 0x5410);   // v0.lo16 | v1.lo16 << 16
  const uint32_t p23 = __byte_perm((uint32_t)v2, (uint32_t)v3, 0x5410);
  lo = __byte_perm(p01, p23, 0x6420);
  hi = __byte_perm(p01, p23, 0x7531);
}

// the same for values that still need their clip to int16 (inverse-transform stages, dequantised coefficients): the
// saturating pair pack does the clip
__device__ __forceinline__ void split4_sat(int v0, int v1, int v2, int v3, uint32_t &lo, uint32_t &hi)
{
  const uint32_t p01 = pack2_sat_s16(v0, v1), p23 = pack2_sat_s16(v2, v3);
  lo = __byte_perm(p01, p23, 0x6420);
  hi = __byte_perm(p01, p23, 0x7531);
}

__host__ __device__ __forceinline__ uint32_t pack4(int b0, int b1, int b2, int b3)
{
  return (uint32_t)(b0 & 0xff) | ((uint32_t)(b1 & 0xff) << 8) | ((uint32_t)(b2 & 0xff) << 16) | ((uint32_t)(b3 & 0xff) << 24);
}

// contraction slot -> column inside a 16-wide group (C-fragment pairs of two adjacent 8-wide tiles)
__host__ __device__ __forceinline__ int pi_col(int t, int e) { return 8 * (e >> 1) + 2 * t + (e & 1); }

}  // namespace uvgb
