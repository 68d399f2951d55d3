// hadamard.cuh - the 8-bit 8x8 Hadamard building block shared by picture.cu and rough.cu.
#pragma once
#include "common.cuh"

namespace uvgb {

// SATD of one 8x8 sub-block from its raw row words (two 32-bit words = 8 pixels per row, for
// the original `wo` and the candidate `wc`), normalised like satd_8x8_subblock_generic
// (picture-generic.c:256-348: sum|coef| with |dc|>>2, then (s+2)>>2). When WITH_SAD, also
// returns the plain SAD of the same 64 pixel pairs (sad_NxN, :1054) from the same registers.
//
// 8-bit fast path: differences are 9-bit and the 8x8 Hadamard grows them by 6 bits, so every
// intermediate fits int16 and two values travel in one 32-bit register in "packed-linear" form
// word = hi*65536 + lo (lo signed, its borrow absorbed by hi). That form is closed under 32-bit
// add/sub: every butterfly is ONE integer instruction for two values. Packing (e_j, e_j+2) of a
// row in one word leaves exactly one butterfly stage (distance 2) inside words; it is folded
// into the final sum through |a+b| + |a-b| = 2*max(|a|,|b|).
//
// The result is invariant under a dyadic shift of the row order (row y -> y ^ s): a Walsh-
// Hadamard transform of dyadically shifted input only flips signs of coefficients. Callers use
// this to read rows in a bank-conflict-free swizzled order.
template <bool WITH_SAD>
__device__ __forceinline__ unsigned satd8_words_u8(const uint32_t (&wo)[8][2], const uint32_t (&wc)[8][2],
                                                   unsigned &sad)
{
  int r[8][4];
  unsigned sd = 0;
#pragma unroll
  for (int y = 0; y < 8; ++y) {
#pragma unroll
    for (int hw = 0; hw < 2; ++hw) {
      if (WITH_SAD) sd = __vsadu4(wo[y][hw], wc[y][hw]) + sd;
      // bytes b0 b1 b2 b3 -> (b0 | b2<<16) and (b1 | b3<<16)
      const int oe = (int)__byte_perm(wo[y][hw], 0, 0x4240), oo = (int)__byte_perm(wo[y][hw], 0, 0x4341);
      const int ce = (int)__byte_perm(wc[y][hw], 0, 0x4240), co = (int)__byte_perm(wc[y][hw], 0, 0x4341);
      r[y][hw * 2 + 0] = oe - ce;   // packed-linear (d_e0, d_e2)
      r[y][hw * 2 + 1] = oo - co;   // (d_e1, d_e3)
    }
    // horizontal butterflies at distance 4 and 1 (word level)
    const int a0 = r[y][0] + r[y][2], a1 = r[y][1] + r[y][3], a2 = r[y][0] - r[y][2], a3 = r[y][1] - r[y][3];
    r[y][0] = a0 + a1; r[y][1] = a0 - a1; r[y][2] = a2 + a3; r[y][3] = a2 - a3;
  }
  // vertical butterflies (distance 4, 2, 1 in y), word level
#pragma unroll
  for (int len = 4; len >= 1; len >>= 1)
#pragma unroll
    for (int i = 0; i < 8; i += len << 1)
#pragma unroll
      for (int j = i; j < i + len; ++j)
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const int a = r[j][x], b = r[j + len][x];
          r[j][x] = a + b; r[j + len][x] = a - b;
        }
  // remaining stage (distance 2, inside words) fused with the abs-sum: each word contributes
  // |lo+hi| + |lo-hi|, two VABSDIFF-accumulate instructions
  unsigned s = 0;
  int dc = 0;
#pragma unroll
  for (int y = 0; y < 8; ++y)
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int wv = r[y][x];
      const int lo = (int)(short)(wv & 0xffff);
      const int hi = (wv - lo) >> 16;
      const int u = lo + hi;
      if (y == 0 && x == 0) dc = u;               // all-plus coefficient
      s = __sad(lo, hi, s);
      s = __sad(u, 0, s);
    }
  s = s - (unsigned)abs(dc) + ((unsigned)abs(dc) >> 2);
  sad = sd;
  return (s + 2) >> 2;                             // :343-345
}

}  // namespace uvgb
