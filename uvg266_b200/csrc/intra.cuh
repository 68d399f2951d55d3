// intra.cuh - the pixel-level intra predictor shared by intra.cu (batched / fused rough-search kernels, drop-in
// leaves) and frame.cu (frame pipeline): VVC angular tables, the per-mode decisions of intra_predict_regular
// (intra.c:660-753) folded into a ModeParam on the host, and pred_pixel = one output pixel of planar / DC /
// angular prediction incl. PDPC (intra-generic.c:55-437).
#pragma once
#include "common.cuh"

namespace uvgb {

// intra-generic.c:73-75 (tables of the VVC angular predictor)
static __constant__ short c_sampledisp[32] = { 0, 1, 2, 3, 4, 6, 8, 10, 12, 14, 16, 18, 20, 23, 26, 29, 32,
  35, 39, 45, 51, 57, 64, 73, 86, 102, 128, 171, 256, 341, 512, 1024 };
static const short h_sampledisp[32] = { 0, 1, 2, 3, 4, 6, 8, 10, 12, 14, 16, 18, 20, 23, 26, 29, 32,
  35, 39, 45, 51, 57, 64, 73, 86, 102, 128, 171, 256, 341, 512, 1024 };
static __constant__ short c_invsampledisp[32] = { 0, 16384, 8192, 5461, 4096, 2731, 2048, 1638, 1365, 1170,
  1024, 910, 819, 712, 630, 565, 512, 468, 420, 364, 321, 287, 256, 224, 191, 161, 128, 96, 64, 48, 32, 16 };
static __constant__ signed char c_pre_scale[32] = { 8, 7, 6, 5, 5, 4, 4, 4, 3, 3, 3, 3, 3, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1,
  0, 0, 0, -1, -1, -2, -3 };
static __constant__ signed char c_dist_thres[8] = { 24, 24, 24, 14, 2, 0, 0, 0 };
static const short h_invsampledisp[32] = { 0, 16384, 8192, 5461, 4096, 2731, 2048, 1638, 1365, 1170,
  1024, 910, 819, 712, 630, 565, 512, 468, 420, 364, 321, 287, 256, 224, 191, 161, 128, 96, 64, 48, 32, 16 };
static const signed char h_pre_scale[32] = { 8, 7, 6, 5, 5, 4, 4, 4, 3, 3, 3, 3, 3, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1,
  0, 0, 0, -1, -1, -2, -3 };
static const signed char h_dist_thres[8] = { 24, 24, 24, 14, 2, 0, 0, 0 };
// 4-tap cubic interpolation filter, one row per 1/32 phase (intra-generic.c:77-111)
static __constant__ signed char c_cubic[32][4] = {
  { 0, 64, 0, 0 }, { -1, 63, 2, 0 }, { -2, 62, 4, 0 }, { -2, 60, 7, -1 }, { -2, 58, 10, -2 }, { -3, 57, 12, -2 },
  { -4, 56, 14, -2 }, { -4, 55, 15, -2 }, { -4, 54, 16, -2 }, { -5, 53, 18, -2 }, { -6, 52, 20, -2 },
  { -6, 49, 24, -3 }, { -6, 46, 28, -4 }, { -5, 44, 29, -4 }, { -4, 42, 30, -4 }, { -4, 39, 33, -4 },
  { -4, 36, 36, -4 }, { -4, 33, 39, -4 }, { -4, 30, 42, -4 }, { -4, 29, 44, -5 }, { -4, 28, 46, -6 },
  { -3, 24, 49, -6 }, { -2, 20, 52, -6 }, { -2, 18, 53, -5 }, { -2, 16, 54, -4 }, { -2, 15, 55, -4 },
  { -2, 14, 56, -4 }, { -2, 12, 57, -3 }, { -2, 10, 58, -2 }, { -1, 7, 60, -2 }, { 0, 4, 62, -2 }, { 0, 2, 63, -1 } };

// Everything about one intra mode that does not depend on the pixel, derived on the host.
struct ModeParam {
  signed char kind;         // 0 planar, 1 DC, 2 angular
  signed char vertical;     // angular: pred_mode >= 34
  signed char filtered;     // use the [1 2 1]-smoothed reference rows
  signed char frac;         // angular: (|sample_disp| & 31) != 0 -> 4-tap interpolation
  signed char cubic;        // angular: cubic (1) or the "gaussian" filter derived from the phase (0)
  signed char pdpc;         // apply PDPC
  signed char scale;        // PDPC scale
  signed char pad;
  short sample_disp;        // angular: displacement per line in 1/32 samples (signed)
  short inv_angle;          // angular: modedisp2invsampledisp
};
constexpr int kMaxModes = 67;
struct ModeTable { ModeParam m[kMaxModes]; int k; };

// intra_predict_regular's decisions for an n x n luma block (square: no wide-angle remap).
static ModeParam make_mode(int n, int mode)
{
  ModeParam p{};
  const int l2 = ilog2_host(n);
  if (mode == 0) {
    p.kind = 0; p.filtered = n * n > 32; p.pdpc = n >= 4; p.scale = (signed char)((2 * l2 - 2) >> 2);
    return p;
  }
  if (mode == 1) {
    p.kind = 1; p.filtered = 0; p.pdpc = n >= 4; p.scale = (signed char)((2 * l2 - 2) >> 2);
    return p;
  }
  p.kind = 2;
  const int pred_mode = mode;
  p.vertical = pred_mode >= 34;
  const int mode_disp = p.vertical ? pred_mode - 50 : 18 - pred_mode;
  const int ad = mode_disp < 0 ? -mode_disp : mode_disp;
  p.sample_disp = (short)((mode_disp < 0 ? -1 : 1) * h_sampledisp[ad]);
  p.inv_angle = h_invsampledisp[ad];
  p.frac = (h_sampledisp[ad] & 31) != 0;
  const int thr = h_dist_thres[l2];                         // (log2w + log2h) >> 1 = log2 n
  int d50 = pred_mode - 50, d18 = pred_mode - 18;
  if (d50 < 0) d50 = -d50;
  if (d18 < 0) d18 = -d18;
  const int dist = d50 < d18 ? d50 : d18;
  p.cubic = !(dist > thr && p.frac);
  p.filtered = (n != 4) && dist > thr && !p.frac;             // intra.c:691-712
  int scale = l2 - h_pre_scale[ad];
  if (scale > 2) scale = 2;
  if (p.sample_disp == 0) {                                  // pure horizontal / vertical (:254-280)
    p.pdpc = n >= 4; p.scale = (signed char)((2 * l2 - 2) >> 2);
  } else {
    p.pdpc = n >= 4 && mode_disp > 0 && scale >= 0;           // :234-243
    p.scale = (signed char)scale;
  }
  return p;
}

__device__ __forceinline__ int ilog2_dev(int v) { return 31 - __clz(v); }

// Reference rows of one block in shared memory: unfiltered and smoothed, top and left, each
// `pitch` pixels long (>= 2n + 1 + 8, tail = last sample replicated).
template <typename Pix>
struct RefRows { const Pix *top, *left, *ftop, *fleft; };

template <typename Pix>
__device__ __forceinline__ int pred_pixel(const ModeParam &mp, const RefRows<Pix> &rr, int n, int l2, int X, int Y,
                                          int dc_val, int pixel_max)
{
  const Pix *top = mp.filtered ? rr.ftop : rr.top, *left = mp.filtered ? rr.fleft : rr.left;
  if (mp.kind != 2) {
    int v;
    if (mp.kind == 0) {                                      // planar :306-360
      const int tr = top[n + 1], bl = left[n + 1];
      const int hor = ((int)left[Y + 1] << l2) + (X + 1) * (tr - (int)left[Y + 1]);
      const int ver = ((int)top[X + 1] << l2) + (Y + 1) * (bl - (int)top[X + 1]);
      v = ((hor << l2) + (ver << l2) + (1 << (2 * l2))) >> (1 + 2 * l2);
    } else {
      v = dc_val;
    }
    if (mp.pdpc) {                                           // uvg_pdpc_planar_dc_generic :414-437
      const int wT = 32 >> min(31, (Y << 1) >> mp.scale), wL = 32 >> min(31, (X << 1) >> mp.scale);
      v = v + ((wL * ((int)left[Y + 1] - v) + wT * ((int)top[X + 1] - v) + 32) >> 6);
      v &= (sizeof(Pix) == 1) ? 0xff : 0xffff;               // stored into uvg_pixel
    }
    return v;
  }
  // angular :55-299: horizontal modes are predicted transposed
  const Pix *mainr = mp.vertical ? top : left, *side = mp.vertical ? left : top;
  const int r = mp.vertical ? Y : X, c = mp.vertical ? X : Y;
  int v;
  if (mp.sample_disp == 0) {
    v = mainr[1 + c];
    if (mp.pdpc && c < min(3 << mp.scale, n)) {
      const int wL = 32 >> ((2 * c) >> mp.scale);
      v = clip3(0, pixel_max, v + ((wL * ((int)side[1 + r] - (int)mainr[0]) + 32) >> 6));
    }
    return v;
  }
  const int delta_pos = (r + 1) * mp.sample_disp;
  const int di = delta_pos >> 5, fr = delta_pos & 31;
  auto MAIN = [&](int i) -> int {
    if (i >= 0) return mainr[i];
    // negative part of the main reference = side reference projected through the inverse angle (:145-156)
    return side[min((-i * mp.inv_angle + 256) >> 9, n)];
  };
  if (mp.frac) {
    const int i0 = di + c;
    int f0, f1, f2, f3;
    if (mp.cubic) { f0 = c_cubic[fr][0]; f1 = c_cubic[fr][1]; f2 = c_cubic[fr][2]; f3 = c_cubic[fr][3]; }
    else { f0 = 16 - (fr >> 1); f1 = 32 - (fr >> 1); f2 = 16 + (fr >> 1); f3 = fr >> 1; }
    v = clip3(0, pixel_max, (f0 * MAIN(i0) + f1 * MAIN(i0 + 1) + f2 * MAIN(i0 + 2) + f3 * MAIN(i0 + 3) + 32) >> 6);
  } else {
    v = MAIN(c + di + 1);
  }
  if (mp.pdpc && c < min(3 << mp.scale, n)) {
    const int inv_sum = 256 + (c + 1) * mp.inv_angle;
    const int wL = 32 >> ((2 * c) >> mp.scale);
    const int l = side[r + (inv_sum >> 9) + 1];
    v = (v + ((wL * (l - v) + 32) >> 6)) & ((sizeof(Pix) == 1) ? 0xff : 0xffff);
  }
  return v;
}

__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c)
{
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}


// Bytes the scratch behind `ext0` must mirror from the main reference row (ext0[0 .. kExtMirror) = mainr[0 .. kExtMirror)):
// a window that starts left of index 0 ends at most 11 bytes right of it, so a row's four words come from ONE base.
constexpr int kExtMirror = 12;

// PDPC of an angular mode does not depend on the row: per column cc the side-reference offset
// ((256 + (cc + 1) * inv_angle) >> 9) + 1 and the weight 32 >> ((2 cc) >> scale). A caller that generates many rows of
// one (mode, strip) computes them once (pdpc_tab) - one byte each, zero weight beyond the PDPC range - and the row
// generator picks them with one PRMT instead of recomputing five instructions per pixel.
struct PdpcTab { uint32_t off[2], w[2]; };
__device__ __forceinline__ PdpcTab pdpc_tab(const ModeParam &mp, int sx, int lim)
{
  PdpcTab t = { { 0, 0 }, { 0, 0 } };
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int cc = 8 * sx + c;
    if (cc < lim) {
      t.off[c >> 2] |= (uint32_t)(((256 + (cc + 1) * mp.inv_angle) >> 9) + 1) << (8 * (c & 3));
      t.w[c >> 2] |= (uint32_t)(32 >> ((2 * cc) >> mp.scale)) << (8 * (c & 3));
    }
  }
  return t;
}

// One row (8 pixels: work-row r, work-columns 8*sx..8*sx+7) of an angular prediction. TAB: PDPC offsets / weights from `tab`.
template <bool TAB = false>
__device__ __forceinline__ void angular_row8(const ModeParam &mp, const uint8_t *mainr, const uint8_t *side,
                                             const uint8_t *ext0, int r, int sx, int lim, uint32_t &lo, uint32_t &hi,
                                             const PdpcTab &tab = PdpcTab{})
{
  const int dpos = (r + 1) * mp.sample_disp, di = dpos >> 5, fr = dpos & 31;
  int cw;
  if (!mp.frac) cw = 64 << 8;                                              // integer slope: copy tap 1
  else if (mp.cubic) cw = *reinterpret_cast<const int *>(c_cubic[fr]);
  else { const int h = fr >> 1; cw = (16 - h) | ((32 - h) << 8) | ((16 + h) << 16) | (h << 24); }
  const int i0 = di + 8 * sx, b0 = i0 & ~3, sh = (i0 & 3) * 8;
  const uint32_t *wp = reinterpret_cast<const uint32_t *>((b0 < 0 ? ext0 : mainr) + b0);
  const uint32_t W0 = wp[0], W1 = wp[1], W2 = wp[2], W3 = wp[3];
  const uint32_t A0 = __funnelshift_r(W0, W1, sh), A1 = __funnelshift_r(W1, W2, sh), A2 = __funnelshift_r(W2, W3, sh);
  int v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t a = c < 4 ? A0 : A1, b = c < 4 ? A1 : A2;
    const uint32_t win = (c & 3) ? __funnelshift_r(a, b, 8 * (c & 3)) : a;
    v[c] = dp4a_us(win, cw, 32) >> 6;                                     // only the cubic taps can leave 0..255
  }
  if (!mp.pdpc) {
    lo = pack4_sat_u8(v[0], v[1], v[2], v[3]);
    hi = pack4_sat_u8(v[4], v[5], v[6], v[7]);
    return;
  }
  if (mp.cubic) {
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = min(max(v[c], 0), 255);
  }
  if (mp.sample_disp == 0) {                                              // :254-280
    const int d = (int)side[1 + r] - (int)mainr[0];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int cc = 8 * sx + c;
      if (TAB) {                                                           // weight 0 beyond the range: (0 * d + 32) >> 6 = 0
        const int w = (int)__byte_perm(tab.w[c >> 2], 0, 0x4440 + (c & 3));
        v[c] = min(max(v[c] + ((w * d + 32) >> 6), 0), 255);
      } else if (cc < lim) {
        v[c] = min(max(v[c] + (((32 >> ((2 * cc) >> mp.scale)) * d + 32) >> 6), 0), 255);
      }
    }
  } else {                                                                 // :234-251
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int cc = 8 * sx + c;
      if (TAB) {
        const int w = (int)__byte_perm(tab.w[c >> 2], 0, 0x4440 + (c & 3));
        const int l = side[r + (int)__byte_perm(tab.off[c >> 2], 0, 0x4440 + (c & 3))];   // offset 0 (any sample) where w = 0
        v[c] = (v[c] + ((w * (l - v[c]) + 32) >> 6)) & 255;
      } else if (cc < lim) {
        const int l = side[r + ((256 + (cc + 1) * mp.inv_angle) >> 9) + 1];
        v[c] = (v[c] + (((32 >> ((2 * cc) >> mp.scale)) * (l - v[c]) + 32) >> 6)) & 255;
      }
    }
  }
  lo = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
  hi = v[4] | (v[5] << 8) | (v[6] << 16) | (v[7] << 24);
}


// One row (8 pixels: row Y, columns 8*sx..8*sx+7) of a planar (kind 0) or DC (kind 1) prediction incl. PDPC
// (intra-generic.c:306-360, :414-437; intra.c:236-273), n >= 4.
__device__ __forceinline__ void planar_dc_row8(const ModeParam &mp, const uint8_t *top, const uint8_t *left, int dc, int n,
                                               int l2, int Y, int sx, uint32_t &lo, uint32_t &hi)
{
  const int tr = top[n + 1], bl = left[n + 1], lY = left[Y + 1];
  const int wT = 32 >> min(31, (Y << 1) >> mp.scale);
  uint32_t w2[2] = { 0, 0 };
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int X = 8 * sx + c, tX = top[X + 1];
    int v = dc;
    if (mp.kind == 0) {
      const int hor = (lY << l2) + (X + 1) * (tr - lY), ver = (tX << l2) + (Y + 1) * (bl - tX);
      v = ((hor << l2) + (ver << l2) + (1 << (2 * l2))) >> (1 + 2 * l2);
    }
    if (mp.pdpc) {
      const int wL = 32 >> min(31, (X << 1) >> mp.scale);
      v = (v + ((wL * (lY - v) + wT * (tX - v) + 32) >> 6)) & 255;
    }
    w2[c >> 2] |= (uint32_t)v << (8 * (c & 3));
  }
  lo = w2[0]; hi = w2[1];
}

// Fused rough search straight from a picture plane (frame.cu): blocks in CTU-major order, originals and reference
// rows read from the plane with border replication instead of from gathered arrays.
struct PlaneSrc { const uint8_t *plane; int stride, W, H, ctu_cols, ctu_begin; };
int launch_intra_rough_plane(int n, const PlaneSrc &src, const int8_t *modes, int k, uint32_t *out, size_t n_blocks,
                             cudaStream_t st);

}  // namespace uvgb
