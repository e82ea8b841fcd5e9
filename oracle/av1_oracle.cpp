// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline leg may load this library.
//
// Scalar CPU restatement of the per-chunk AV1 encode path that the reference hands to av1an+SVT-AV1
// (/root/reference/crates/daemon/src/encode/av1an.rs:79-139).  The arithmetic of that path lives in
// third-party code that is not under /root/reference (Docker image masterofzen/av1an:master, no pinned
// version: scripts/av1an-docker:11), so the normative parts restate the AV1 specification
//   7.11.2 intra prediction, 7.12.3 dequantisation, 7.13 inverse transforms + reconstruction,
//   7.14 loop filter, 7.15 CDEF, 7.17 loop restoration
// and are pinned (tests/test_oracle_*.py) against libaom 3.13.1's own C reference functions
// (av1_inv_txfm2d_add_*_c, aom_highbd_*_predictor_*_c, aom_highbd_lpf_*_c, cdef_filter_*_c, ...) and
// against the dav1d 1.5.3 and libaom decoders on complete bitstreams.
// The encoder-side (non-normative) parts -- forward transform, quantiser, mode decision -- are the
// definition the CUDA path must reproduce bit for bit.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "../av1_base_b200/csrc/av1b_types.h"
#include "../av1_base_b200/csrc/av1_inv_txfm1d.h"   // normative butterfly graphs (pinned vs libaom av1_idct*)
#include "../av1_base_b200/csrc/av1_tables.h"         // normative constant tables (extracted from the libaom binary): one copy
#include "../av1_base_b200/csrc/av1_fwd_matrices.h"   // generated from the normative inverse transforms
#include "../av1_base_b200/csrc/av1_qm_tables.h"      // Quantizer_Matrix of the square sizes (same origin; the decoders pin it)

using namespace av1tx;

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int ilog2i(int n) { int k = 0; while ((1 << k) < n) k++; return k; }

// ------------------------------------------------------------------------------------------------
// transforms
// ------------------------------------------------------------------------------------------------
enum { T_DCT = 0, T_ADST = 1, T_FLIP = 2, T_IDT = 3 };
static const uint8_t kVType[16] = {T_DCT, T_ADST, T_DCT, T_ADST, T_FLIP, T_DCT, T_FLIP, T_ADST, T_FLIP,
                                   T_IDT, T_DCT, T_IDT, T_ADST, T_IDT, T_FLIP, T_IDT};
static const uint8_t kHType[16] = {T_DCT, T_DCT, T_ADST, T_ADST, T_DCT, T_FLIP, T_FLIP, T_FLIP, T_ADST,
                                   T_IDT, T_IDT, T_DCT, T_IDT, T_ADST, T_IDT, T_FLIP};

// Transform_Row_Shift (spec 7.13.3), indexed [log2w-2][log2h-2]
static const int8_t kRowShift[5][5] = {
    // h: 4  8  16 32 64
    {0, 0, 1, -1, -1},   // w = 4
    {0, 1, 1, 2, -1},    // w = 8
    {1, 1, 2, 1, 2},     // w = 16
    {-1, 2, 1, 2, 1},    // w = 32
    {-1, -1, 2, 1, 2},   // w = 64
};

static void inv1d(int t, int n, int32_t* x, int range) {
  if (t == T_DCT) {
    switch (n) { case 4: idct<4>(x, range); break; case 8: idct<8>(x, range); break;
      case 16: idct<16>(x, range); break; case 32: idct<32>(x, range); break; default: idct<64>(x, range); }
  } else if (t == T_IDT) {
    switch (n) { case 4: iidentity<4>(x, range); break; case 8: iidentity<8>(x, range); break;
      case 16: iidentity<16>(x, range); break; default: iidentity<32>(x, range); }
  } else {
    switch (n) { case 4: iadst4(x, range); break; case 8: iadst8(x, range); break; default: iadst16(x, range); }
  }
}

// 2-D inverse transform + reconstruction (spec 7.13.3 / libaom inv_txfm2d_add_c).
// dq: dequantised coefficients, SPEC layout (row = vertical frequency), stride cstride, only the
// top-left min(w,32) x min(h,32) are read.
extern "C" void orc_inv_txfm2d_add(const int32_t* dq, int cstride, uint16_t* dst, int dstride, int w, int h,
                                   int tx_type, int bd) {
  const int lw = ilog2i(w), lh = ilog2i(h);
  const int row_shift = kRowShift[lw - 2][lh - 2];
  const int vt = kVType[tx_type], ht = kHType[tx_type];
  const int row_range = bd + 8, col_range = std::max(bd + 6, 16);
  const bool rect = abs(lw - lh) == 1;
  const int cw = std::min(w, 32), ch = std::min(h, 32);
  std::vector<int32_t> buf((size_t)w * h);
  int32_t T[64];
  for (int i = 0; i < h; i++) {
    for (int j = 0; j < w; j++) {
      int32_t v = (i < ch && j < cw) ? dq[i * cstride + j] : 0;
      if (rect) v = (int32_t)(((int64_t)v * 2896 + 2048) >> 12);
      T[j] = sat(v, row_range);
    }
    inv1d(ht, w, T, row_range);
    for (int j = 0; j < w; j++) {
      int32_t v = T[ht == T_FLIP ? w - 1 - j : j];
      if (row_shift > 0) v = (v + (1 << (row_shift - 1))) >> row_shift;
      buf[(size_t)i * w + j] = v;
    }
  }
  const int maxv = (1 << bd) - 1;
  for (int j = 0; j < w; j++) {
    for (int i = 0; i < h; i++) T[i] = sat(buf[(size_t)i * w + j], col_range);
    inv1d(vt, h, T, col_range);
    for (int i = 0; i < h; i++) {
      int32_t v = T[vt == T_FLIP ? h - 1 - i : i];
      v = (v + 8) >> 4;
      uint16_t* p = dst + (size_t)i * dstride + j;
      *p = (uint16_t)clampi((int)*p + v, 0, maxv);
    }
  }
}

// n blocks in a row: block b reads dq + b * min(w,32) * min(h,32) and updates dst + b * w * h (the 4096-block suite).
extern "C" void orc_inv_txfm2d_add_batch(int n, const int32_t* dq, uint16_t* dst, int w, int h, int tx_type, int bd) {
  const int cw = std::min(w, 32), ch = std::min(h, 32);
  for (int b = 0; b < n; b++) orc_inv_txfm2d_add(dq + (size_t)b * cw * ch, cw, dst + (size_t)b * w * h, w, w, h, tx_type, bd);
}

static inline int fwd_coef(int t, int n, int k, int i) {
  if (t == T_IDT) {
    if (k != i) return 0;
    return n == 4 ? 5793 : n == 8 ? 8192 : n == 16 ? 11585 : 16384;
  }
  if (t == T_DCT) {
    switch (n) { case 4: return av1t_fwd_dct4[k][i]; case 8: return av1t_fwd_dct8[k][i];
      case 16: return av1t_fwd_dct16[k][i]; case 32: return av1t_fwd_dct32[k][i]; default: return av1t_fwd_dct64[k][i]; }
  }
  switch (n) { case 4: return av1t_fwd_adst4[k][i]; case 8: return av1t_fwd_adst8[k][i]; default: return av1t_fwd_adst16[k][i]; }
}

// Encoder-side forward transform (our own definition, matrix form, exact integer arithmetic):
//   x' = resid << 2 (flipped for FLIPADST), t = (Fv x' + 2^11) >> 12 (columns),
//   acc = Fh t (rows, 64-bit), coef = (acc * mul + rnd) >> (24 + log2(w*h) - rowShift - 4)
// with mul = 4096, or 5793 for 2:1 rectangles.  Output: SPEC layout, stride cstride, only
// min(w,32) x min(h,32) coefficients are produced (AV1 zero-out of 64-point transforms).
extern "C" void orc_fwd_txfm2d(const int16_t* resid, int rstride, int32_t* coef, int cstride, int w, int h,
                               int tx_type) {
  const int lw = ilog2i(w), lh = ilog2i(h);
  const int vt = kVType[tx_type], ht = kHType[tx_type];
  const int cw = std::min(w, 32), ch = std::min(h, 32);
  const bool rect = abs(lw - lh) == 1;
  const int sh = 24 + lw + lh - kRowShift[lw - 2][lh - 2] - 4;
  const int64_t mul = rect ? 5793 : 4096;
  std::vector<int32_t> t((size_t)ch * w);
  for (int k = 0; k < ch; k++)
    for (int j = 0; j < w; j++) {
      int32_t acc = 0;
      const int jj = ht == T_FLIP ? w - 1 - j : j;
      for (int i = 0; i < h; i++) {
        const int ii = vt == T_FLIP ? h - 1 - i : i;
        acc += fwd_coef(vt, h, k, i) * ((int32_t)resid[ii * rstride + jj] * 4);
      }
      t[(size_t)k * w + j] = (acc + 2048) >> 12;
    }
  for (int k = 0; k < ch; k++)
    for (int l = 0; l < cw; l++) {
      int64_t acc = 0;
      for (int j = 0; j < w; j++) acc += (int64_t)fwd_coef(ht, w, l, j) * t[(size_t)k * w + j];
      coef[k * cstride + l] = (int32_t)((acc * mul + ((int64_t)1 << (sh - 1))) >> sh);
    }
}

// n blocks in a row (the forward-transform suite): block b reads resid + b*w*h, writes coef + b * min(w,32) * min(h,32)
extern "C" void orc_fwd_txfm2d_batch(int n, const int16_t* resid, int32_t* coef, int w, int h, int tx_type) {
  const int cw = std::min(w, 32), ch = std::min(h, 32);
  for (int b = 0; b < n; b++) orc_fwd_txfm2d(resid + (size_t)b * w * h, w, coef + (size_t)b * cw * ch, cw, w, h, tx_type);
}

// ------------------------------------------------------------------------------------------------
// quantiser
// ------------------------------------------------------------------------------------------------
static inline int dc_q(int qidx, int bd) { return bd == 8 ? av1t_dc_q_8[qidx] : bd == 10 ? av1t_dc_q_10[qidx] : av1t_dc_q_12[qidx]; }
static inline int ac_q(int qidx, int bd) { return bd == 8 ? av1t_ac_q_8[qidx] : bd == 10 ? av1t_ac_q_10[qidx] : av1t_ac_q_12[qidx]; }
static inline int tx_scale_shift(int w, int h) { const int p = w * h; return (p > 256) + (p > 1024); }

// Quantisation matrices (spec 7.12.3, using_qmatrix = 1): the step of coefficient (i, j) becomes
//   q2 = Round2(q * Quantizer_Matrix[level][plane > 0][offset(size) + i * tw + j], 5)
// for levels 0..14 (15 = flat).  The frame's levels are oracle state set by orc_set_qm (tests / chain.py) so that the
// frame functions keep their signatures; product: Av1bFrameParams.qm_level -> IntraLaunch / InterLaunch.qm.
static int g_qm_level[2] = {15, 15};
extern "C" void orc_set_qm(int level_y, int level_uv) { g_qm_level[0] = level_y; g_qm_level[1] = level_uv; }
// the level SVT-AV1 / libaom derive from the quantiser index (aom_get_qmlevel): first + qindex * (last + 1 - first) / 256
extern "C" int orc_qm_level(int qidx, int first, int last) { return first + (qidx * (last + 1 - first)) / 256; }
static inline int qm_step(int q, int plane, int cw, int i, int j) {
  const int lvl = g_qm_level[plane > 0];
  if (lvl >= 15) return q;
  return (q * av1t_qm_sq[lvl][plane > 0][av1t_qm_sq_offset(cw) + i * cw + j] + 16) >> 5;
}

// Encoder quantiser (ours): level = min((|c| << s) + ((dqv * rnd) >> 7)) / dqv, 32767)
// Dequantiser (normative, spec 7.12.3): dq = ((level * dqv) & 0xFFFFFF) >> s, clipped to bd+8 bits.
// dqv = the step of the position (dc / ac quantiser, weighted by the quantisation matrix when one is in force).
// Returns eob-independent stats; levels/dq in SPEC layout.  Square 2-D transforms only when a matrix is in force.
extern "C" void orc_quant_dequant(const int32_t* coef, int cstride, int16_t* lev, int lstride, int32_t* dq,
                                  int dstride, int w, int h, int qidx, int bd, int rnd, int plane) {
  const int cw = std::min(w, 32), ch = std::min(h, 32), s = tx_scale_shift(w, h);
  const int lim = (1 << (7 + bd)) - 1;
  for (int i = 0; i < ch; i++)
    for (int j = 0; j < cw; j++) {
      const int dqv = qm_step((i | j) ? ac_q(qidx, bd) : dc_q(qidx, bd), plane, cw, i, j);
      const int32_t c = coef[i * cstride + j];
      const int64_t a = (int64_t)(c < 0 ? -(int64_t)c : c) << s;
      int64_t l = (a + ((dqv * rnd) >> 7)) / dqv;
      if (l > 32767) l = 32767;
      int64_t d = ((l * dqv) & 0xFFFFFF) >> s;
      if (d > lim) d = lim;   // symmetric here because the magnitude is clipped first;
      lev[i * lstride + j] = (int16_t)(c < 0 ? -l : l);
      int32_t dv = (int32_t)d;
      if (c < 0) { dv = -dv; if (dv < -lim - 1) dv = -lim - 1; }
      dq[i * dstride + j] = dv;
    }
}

extern "C" void orc_dequant(const int16_t* lev, int lstride, int32_t* dq, int dstride, int w, int h, int qidx, int bd, int plane) {
  const int cw = std::min(w, 32), ch = std::min(h, 32), s = tx_scale_shift(w, h);
  const int lim = (1 << (7 + bd)) - 1;
  for (int i = 0; i < ch; i++)
    for (int j = 0; j < cw; j++) {
      const int dqv = qm_step((i | j) ? ac_q(qidx, bd) : dc_q(qidx, bd), plane, cw, i, j);
      const int l = lev[i * lstride + j];
      const int64_t a = l < 0 ? -l : l;
      int64_t d = ((a * dqv) & 0xFFFFFF) >> s;
      if (l < 0) d = -d;
      dq[i * dstride + j] = (int32_t)std::max<int64_t>(-lim - 1, std::min<int64_t>(lim, d));
    }
}

// ------------------------------------------------------------------------------------------------
// intra prediction (spec 7.11.2, enable_intra_edge_filter = 0: no edge filter / upsampling)
// ------------------------------------------------------------------------------------------------
// above[-1 .. w+h-1], left[-1 .. w+h-1] are prepared by the caller.
extern "C" void orc_intra_predict(uint16_t* dst, int dstride, int w, int h, const uint16_t* above,
                                  const uint16_t* left, int mode, int angle_delta, int have_above, int have_left,
                                  int bd) {
  const int lw = ilog2i(w), lh = ilog2i(h);
  auto P = [&](int i, int j) -> uint16_t& { return dst[i * dstride + j]; };
  if (mode == AV1B_DC_PRED) {
    int v;
    if (have_above && have_left) {
      int s = 0;
      for (int k = 0; k < w; k++) s += above[k];
      for (int k = 0; k < h; k++) s += left[k];
      s += (w + h) >> 1;
      v = s / (w + h);
    } else if (have_left) {
      int s = 0; for (int k = 0; k < h; k++) s += left[k];
      v = (s + (h >> 1)) >> lh;
    } else if (have_above) {
      int s = 0; for (int k = 0; k < w; k++) s += above[k];
      v = (s + (w >> 1)) >> lw;
    } else v = 1 << (bd - 1);
    for (int i = 0; i < h; i++) for (int j = 0; j < w; j++) P(i, j) = (uint16_t)v;
    return;
  }
  if (mode == AV1B_PAETH_PRED) {
    const int tl = above[-1];
    for (int i = 0; i < h; i++) for (int j = 0; j < w; j++) {
      const int base = above[j] + left[i] - tl;
      const int pl = abs(base - left[i]), pt = abs(base - above[j]), ptl = abs(base - tl);
      P(i, j) = (pl <= pt && pl <= ptl) ? left[i] : (pt <= ptl ? above[j] : (uint16_t)tl);
    }
    return;
  }
  if (mode == AV1B_SMOOTH_PRED || mode == AV1B_SMOOTH_V_PRED || mode == AV1B_SMOOTH_H_PRED) {
    const uint8_t* wy = av1t_smooth_weights + h - 4;
    const uint8_t* wx = av1t_smooth_weights + w - 4;
    for (int i = 0; i < h; i++) for (int j = 0; j < w; j++) {
      if (mode == AV1B_SMOOTH_PRED) {
        int s = wy[i] * above[j] + (256 - wy[i]) * left[h - 1] + wx[j] * left[i] + (256 - wx[j]) * above[w - 1];
        P(i, j) = (uint16_t)((s + 256) >> 9);
      } else if (mode == AV1B_SMOOTH_V_PRED) {
        int s = wy[i] * above[j] + (256 - wy[i]) * left[h - 1];
        P(i, j) = (uint16_t)((s + 128) >> 8);
      } else {
        int s = wx[j] * left[i] + (256 - wx[j]) * above[w - 1];
        P(i, j) = (uint16_t)((s + 128) >> 8);
      }
    }
    return;
  }
  // directional (V, H and the six diagonal modes share the angle machinery)
  const int angle = av1t_mode_to_angle[mode] + 3 * angle_delta;
  if (angle == 90) { for (int i = 0; i < h; i++) for (int j = 0; j < w; j++) P(i, j) = above[j]; return; }
  if (angle == 180) { for (int i = 0; i < h; i++) for (int j = 0; j < w; j++) P(i, j) = left[i]; return; }
  if (angle < 90) {
    const int dx = av1t_dr_intra_derivative[angle];
    const int max_base = w + h - 1;
    for (int i = 0; i < h; i++) for (int j = 0; j < w; j++) {
      const int idx = (i + 1) * dx, base = (idx >> 6) + j, sh = (idx >> 1) & 31;
      P(i, j) = base < max_base ? (uint16_t)((above[base] * (32 - sh) + above[base + 1] * sh + 16) >> 5) : above[max_base];
    }
  } else if (angle < 180) {
    const int dx = av1t_dr_intra_derivative[180 - angle], dy = av1t_dr_intra_derivative[angle - 90];
    for (int i = 0; i < h; i++) for (int j = 0; j < w; j++) {
      int idx = (j << 6) - (i + 1) * dx;
      int base = idx >> 6;
      if (base >= -1) {
        const int sh = (idx >> 1) & 31;
        P(i, j) = (uint16_t)((above[base] * (32 - sh) + above[base + 1] * sh + 16) >> 5);
      } else {
        idx = (i << 6) - (j + 1) * dy;
        base = idx >> 6;
        const int sh = (idx >> 1) & 31;
        P(i, j) = (uint16_t)((left[base] * (32 - sh) + left[base + 1] * sh + 16) >> 5);
      }
    }
  } else {
    const int dy = av1t_dr_intra_derivative[270 - angle];
    const int max_base = w + h - 1;
    for (int i = 0; i < h; i++) for (int j = 0; j < w; j++) {
      const int idx = (j + 1) * dy, base = (idx >> 6) + i, sh = (idx >> 1) & 31;
      P(i, j) = base < max_base ? (uint16_t)((left[base] * (32 - sh) + left[base + 1] * sh + 16) >> 5) : left[max_base];
    }
  }
}

// Edge preparation (spec 7.11.2 steps 1-4).  rec: reconstructed plane; (x,y) sample position of the
// transform block; max_x / max_y: last valid sample of the plane.  Writes above[-1..w+h-1] and
// left[-1..w+h-1] (buffers must have one element of headroom before index 0).
static void build_edges(const uint16_t* rec, int stride, int x, int y, int w, int h, int have_above,
                        int have_left, int have_above_right, int have_below_left, int max_x, int max_y, int bd,
                        uint16_t* above, uint16_t* left) {
  const int n = w + h;
  const int base = 1 << (bd - 1);
  if (have_above) {
    const uint16_t* r = rec + (size_t)(y - 1) * stride;
    for (int i = 0; i < w; i++) above[i] = r[std::min(max_x, x + i)];
    for (int i = w; i < n; i++) above[i] = have_above_right ? r[std::min(max_x, x + i)] : r[std::min(max_x, x + w - 1)];
  } else {
    const uint16_t v = have_left ? rec[(size_t)y * stride + x - 1] : (uint16_t)(base - 1);
    for (int i = 0; i < n; i++) above[i] = v;
  }
  if (have_left) {
    for (int i = 0; i < h; i++) left[i] = rec[(size_t)std::min(max_y, y + i) * stride + x - 1];
    for (int i = h; i < n; i++)
      left[i] = have_below_left ? rec[(size_t)std::min(max_y, y + i) * stride + x - 1]
                                : rec[(size_t)std::min(max_y, y + h - 1) * stride + x - 1];
  } else {
    const uint16_t v = have_above ? rec[(size_t)(y - 1) * stride + x] : (uint16_t)(base + 1);
    for (int i = 0; i < n; i++) left[i] = v;
  }
  uint16_t tl;
  if (have_above && have_left) tl = rec[(size_t)(y - 1) * stride + x - 1];
  else if (have_above) tl = rec[(size_t)(y - 1) * stride + x];
  else if (have_left) tl = rec[(size_t)y * stride + x - 1];
  else tl = (uint16_t)base;
  above[-1] = tl;
  left[-1] = tl;
}

// ------------------------------------------------------------------------------------------------
// cost: sum of absolute 4x4 Hadamard coefficients of (src - pred)
// ------------------------------------------------------------------------------------------------
static int satd4x4(const uint16_t* a, int as, const uint16_t* b, int bs) {
  int d[16];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) d[i * 4 + j] = (int)a[i * as + j] - (int)b[i * bs + j];
  for (int i = 0; i < 4; i++) {   // rows
    int* r = d + i * 4;
    int s0 = r[0] + r[1], s1 = r[0] - r[1], s2 = r[2] + r[3], s3 = r[2] - r[3];
    r[0] = s0 + s2; r[1] = s1 + s3; r[2] = s0 - s2; r[3] = s1 - s3;
  }
  int sum = 0;
  for (int j = 0; j < 4; j++) {   // columns
    int s0 = d[j] + d[4 + j], s1 = d[j] - d[4 + j], s2 = d[8 + j] + d[12 + j], s3 = d[8 + j] - d[12 + j];
    sum += abs(s0 + s2) + abs(s1 + s3) + abs(s0 - s2) + abs(s1 - s3);
  }
  return sum;
}
extern "C" int orc_satd(const uint16_t* a, int as, const uint16_t* b, int bs, int w, int h) {
  int s = 0;
  for (int i = 0; i < h; i += 4) for (int j = 0; j < w; j += 4) s += satd4x4(a + i * as + j, as, b + i * bs + j, bs);
  return s;
}

// ------------------------------------------------------------------------------------------------
// intra frame encode (decisions + reconstruction), tile by tile, superblock by superblock
// ------------------------------------------------------------------------------------------------
static const uint8_t kModeToTxfm[14] = {AV1B_DCT_DCT, AV1B_ADST_DCT, AV1B_DCT_ADST, AV1B_DCT_DCT, AV1B_ADST_ADST,
                                        AV1B_ADST_DCT, AV1B_DCT_ADST, AV1B_DCT_ADST, AV1B_ADST_DCT, AV1B_ADST_ADST,
                                        AV1B_ADST_DCT, AV1B_DCT_ADST, AV1B_ADST_ADST, AV1B_DCT_DCT};

// Fixed-size square partition with forced splits at the frame edge (a block must lie inside the
// mode-info area to be a leaf).
extern "C" void orc_partition_fixed(const Av1bGeom* g, int blk_log2, uint8_t* map /*[h8*w8]*/) {
  for (int y = 0; y < g->h8; y++)
    for (int x = 0; x < g->w8; x++) {
      int bl = blk_log2;
      while (bl > 3) {
        const int n8 = 1 << (bl - 3), x0 = x & ~(n8 - 1), y0 = y & ~(n8 - 1);
        if (x0 + n8 <= g->w8 && y0 + n8 <= g->h8) break;
        bl--;
      }
      map[y * g->w8 + x] = (uint8_t)bl;
    }
}

// Key-frame partition by smoothness (encoder side, ours).  Sums of the 4x4 luma boxes of a 64x64 (then 32x32) block
// are compared with the plane through the block's mean whose slopes come from the half sums (right - left, bottom -
// top): when no box deviates from that plane by more than thr (in box-sum units) the block is coded whole (one
// prediction, one transform); otherwise it splits, down to the fixed 16x16 blocks.  Blocks that do not lie inside the
// picture keep the fixed partition.  Scaled by n^3 (n boxes per side) everything is an integer:
//   n^3 * plane(i, j) = n * S + (2j - n + 1) * 2 (SR - SL) + (2i - n + 1) * 2 (SB - ST).
static bool block_is_smooth(const uint16_t* src_y, int stride, int x0, int y0, int N, int thr) {
  const int n = N / 4;
  int32_t box[16][16];
  int64_t S = 0, SL = 0, SR = 0, ST = 0, SB = 0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      int32_t s = 0;
      for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) s += src_y[(size_t)(y0 + 4 * i + y) * stride + x0 + 4 * j + x];
      box[i][j] = s;
      S += s;
      (j < n / 2 ? SL : SR) += s;
      (i < n / 2 ? ST : SB) += s;
    }
  const int64_t gx = 2 * (SR - SL), gy = 2 * (SB - ST), n3 = (int64_t)n * n * n;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      const int64_t d = n3 * box[i][j] - (n * S + (2 * j - n + 1) * gx + (2 * i - n + 1) * gy);
      if ((d < 0 ? -d : d) > (int64_t)thr * n3) return false;
    }
  return true;
}

extern "C" void orc_partition_smooth(const Av1bGeom* g, const uint16_t* src_y, int stride, int thr, uint8_t* map /*[h8*w8]*/) {
  orc_partition_fixed(g, 4, map);
  for (int y0 = 0; y0 + 64 <= g->height; y0 += 64)
    for (int x0 = 0; x0 + 64 <= g->width; x0 += 64) {
      if (block_is_smooth(src_y, stride, x0, y0, 64, thr)) {
        for (int yy = 0; yy < 8; yy++) for (int xx = 0; xx < 8; xx++) map[(y0 / 8 + yy) * g->w8 + x0 / 8 + xx] = 6;
        continue;
      }
    }
  for (int y0 = 0; y0 + 32 <= g->height; y0 += 32)
    for (int x0 = 0; x0 + 32 <= g->width; x0 += 32) {
      if (map[(y0 / 8) * g->w8 + x0 / 8] == 6) continue;
      if (block_is_smooth(src_y, stride, x0, y0, 32, thr))
        for (int yy = 0; yy < 4; yy++) for (int xx = 0; xx < 4; xx++) map[(y0 / 8 + yy) * g->w8 + x0 / 8 + xx] = 5;
    }
}

struct IntraEnc {
  const Av1bGeom* g;
  int bd, qidx, rnd;
  const uint16_t* src[3];
  int sstride[3];
  uint16_t* rec[3];     // padded planes, stride g->stride[p]
  Av1bBlockInfo* blocks;
  int16_t* coef[3];     // stride g->stride[p]
  // BlockDecoded flags for the current superblock: [plane][y+1][x+1], 4x4 units of the plane
  uint8_t decoded[3][19][19];
  int mi_row_end, mi_col_end, mi_row_start, mi_col_start;   // current tile

  void clear_block_decoded(int r, int c) {
    for (int p = 0; p < 3; p++) {
      const int ss = p > 0;
      const int sbw4 = (mi_col_end - c) >> ss, sbh4 = (mi_row_end - r) >> ss, n = 16 >> ss;
      for (int y = -1; y <= n; y++)
        for (int x = -1; x <= n; x++) {
          uint8_t v = 0;
          if (y < 0 && x < sbw4) v = 1;
          else if (x < 0 && y < sbh4) v = 1;
          decoded[p][y + 1][x + 1] = v;
        }
      decoded[p][n + 1][0] = 0;
    }
  }

  // from_source = true: edges taken from the SOURCE picture (open-loop mode decision);
  // false: from the reconstruction (what the decoder predicts from).
  void edges_for(int p, int mi_r, int mi_c, int n, int sb_r, int sb_c, uint16_t* above, uint16_t* left,
                 int* have_above, int* have_left, bool from_source) {
    const int ss = p > 0;
    const int x = (mi_c * 4) >> ss, y = (mi_r * 4) >> ss;
    const int ha = mi_r > mi_row_start, hl = mi_c > mi_col_start;
    const int x4 = ((mi_c - sb_c) >> ss), y4 = ((mi_r - sb_r) >> ss), n4 = n >> 2;
    const int har = decoded[p][y4 - 1 + 1][x4 + n4 + 1];
    const int hbl = decoded[p][y4 + n4 + 1][x4 - 1 + 1];
    const int max_x = ((g->mi_cols * 4) >> ss) - 1, max_y = ((g->mi_rows * 4) >> ss) - 1;
    if (from_source) build_edges(src[p], sstride[p], x, y, n, n, ha, hl, har, hbl, max_x, max_y, bd, above, left);
    else build_edges(rec[p], g->stride[p], x, y, n, n, ha, hl, har, hbl, max_x, max_y, bd, above, left);
    *have_above = ha; *have_left = hl;
  }

  void block(int mi_r, int mi_c, int bl, int sb_r, int sb_c) {
    static const uint8_t cand[13] = {AV1B_DC_PRED, AV1B_V_PRED, AV1B_H_PRED, AV1B_PAETH_PRED, AV1B_SMOOTH_PRED,
                                     AV1B_SMOOTH_V_PRED, AV1B_SMOOTH_H_PRED, AV1B_D45_PRED, AV1B_D135_PRED,
                                     AV1B_D113_PRED, AV1B_D157_PRED, AV1B_D203_PRED, AV1B_D67_PRED};
    Av1bBlockInfo info;
    memset(&info, 0, sizeof(info));
    info.blk_log2 = (uint8_t)bl;
    uint16_t edge_a[3][130], edge_l[3][130];
    uint16_t pred[64 * 64];
    int16_t resid[64 * 64];
    int32_t cf[32 * 32], dq[32 * 32];
    int16_t lv[32 * 32];
    for (int pass = 0; pass < 2; pass++) {   // 0: luma, 1: chroma (U and V share the mode)
      const int p0 = pass ? 1 : 0, p1 = pass ? 2 : 0;
      const int ss = pass;
      const int n = std::min(1 << (bl - ss), pass ? 32 : 64);
      int ha = 0, hl = 0;
      // mode decision (open loop): edges from the SOURCE picture, minimum SATD over the candidate list, ties ->
      // first in list.  Decisions therefore do not depend on the reconstruction and can be taken for all
      // blocks of a frame in parallel.
      for (int p = p0; p <= p1; p++) edges_for(p, mi_r, mi_c, n, sb_r, sb_c, edge_a[p] + 1, edge_l[p] + 1, &ha, &hl, true);
      int best = -1; int64_t best_cost = 0;
      for (int k = 0; k < 13; k++) {
        int64_t cost = 0;
        for (int p = p0; p <= p1; p++) {
          orc_intra_predict(pred, n, n, n, edge_a[p] + 1, edge_l[p] + 1, cand[k], 0, ha, hl, bd);
          const int x = (mi_c * 4) >> ss, y = (mi_r * 4) >> ss;
          cost += orc_satd(src[p] + (size_t)y * sstride[p] + x, sstride[p], pred, n, n, n);
        }
        if (best < 0 || cost < best_cost) { best = cand[k]; best_cost = cost; }
      }
      if (pass == 0) info.y_mode = (uint8_t)best; else info.uv_mode = (uint8_t)best;
      // closed loop from here: the prediction that is coded uses the reconstructed neighbours
      for (int p = p0; p <= p1; p++) edges_for(p, mi_r, mi_c, n, sb_r, sb_c, edge_a[p] + 1, edge_l[p] + 1, &ha, &hl, false);
      for (int p = p0; p <= p1; p++) {
        const int x = (mi_c * 4) >> ss, y = (mi_r * 4) >> ss;
        int tx_type = AV1B_DCT_DCT;
        if (p > 0) {
          tx_type = kModeToTxfm[best];
          // transform set of the chroma transform size: 32 -> DCT only, 16 -> no 1-D identity types
          if (n >= 32) tx_type = AV1B_DCT_DCT;
        } else {
          info.tx_type_y = AV1B_DCT_DCT;
        }
        orc_intra_predict(pred, n, n, n, edge_a[p] + 1, edge_l[p] + 1, best, 0, ha, hl, bd);
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++)
          resid[i * n + j] = (int16_t)((int)src[p][(size_t)(y + i) * sstride[p] + x + j] - (int)pred[i * n + j]);
        const int cn = std::min(n, 32);
        orc_fwd_txfm2d(resid, n, cf, cn, n, n, tx_type);
        orc_quant_dequant(cf, cn, lv, cn, dq, cn, n, n, qidx, bd, rnd, p);
        // eob = 1 + last non-zero position in scan order
        const int16_t* scan = cn == 4 ? av1t_scan_default_4x4 : cn == 8 ? av1t_scan_default_8x8
                              : cn == 16 ? av1t_scan_default_16x16 : av1t_scan_default_32x32;
        int eob = 0;
        for (int k = cn * cn - 1; k >= 0; k--) if (lv[scan[k]]) { eob = k + 1; break; }
        info.eob[p] = (uint16_t)eob;
        int16_t* cdst = coef[p] + av1b_coef_offset(g->sb_cols, p, x, y);
        for (int i = 0; i < cn * cn; i++) cdst[i] = lv[i];
        uint16_t* rdst = rec[p] + (size_t)y * g->stride[p] + x;
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) rdst[(size_t)i * g->stride[p] + j] = pred[i * n + j];
        if (eob > 0) orc_inv_txfm2d_add(dq, cn, rdst, g->stride[p], n, n, tx_type, bd);
      }
    }
    info.skip = (info.eob[0] == 0 && info.eob[1] == 0 && info.eob[2] == 0);
    // publish side info on every 8x8 unit of the block, and mark the block decoded
    const int n8 = 1 << (bl - 3);
    for (int yy = 0; yy < n8; yy++) for (int xx = 0; xx < n8; xx++)
      blocks[((mi_r >> 1) + yy) * g->w8 + (mi_c >> 1) + xx] = info;
    for (int p = 0; p < 3; p++) {
      const int ss = p > 0, n4 = (1 << (bl - 2)) >> ss;
      const int x4 = (mi_c - sb_c) >> ss, y4 = (mi_r - sb_r) >> ss;
      for (int yy = 0; yy < n4; yy++) for (int xx = 0; xx < n4; xx++) decoded[p][y4 + yy + 1][x4 + xx + 1] = 1;
    }
  }

  void partition(int mi_r, int mi_c, int bl, int sb_r, int sb_c, const uint8_t* map) {
    if (mi_r >= g->mi_rows || mi_c >= g->mi_cols) return;
    const int want = map[(mi_r >> 1) * g->w8 + (mi_c >> 1)];
    if (want >= bl) { block(mi_r, mi_c, bl, sb_r, sb_c); return; }
    const int h = 1 << (bl - 3);   // half size in mi units
    partition(mi_r, mi_c, bl - 1, sb_r, sb_c, map);
    partition(mi_r, mi_c + h, bl - 1, sb_r, sb_c, map);
    partition(mi_r + h, mi_c, bl - 1, sb_r, sb_c, map);
    partition(mi_r + h, mi_c + h, bl - 1, sb_r, sb_c, map);
  }
};

// src: 4:2:0 planes (uint16 samples); rec/coef: padded planes with the geometry's strides;
// part_map: [h8*w8] block log2 sizes (a consistent quadtree, see orc_partition_fixed).
extern "C" int orc_encode_intra_frame(const Av1bGeom* g, int bit_depth, int base_q_idx, int quant_rnd,
                                      const uint16_t* src_y, const uint16_t* src_u, const uint16_t* src_v,
                                      int sy_stride, int suv_stride, const uint8_t* part_map,
                                      uint16_t* rec_y, uint16_t* rec_u, uint16_t* rec_v, Av1bBlockInfo* blocks,
                                      int16_t* coef_y, int16_t* coef_u, int16_t* coef_v) {
  IntraEnc e;
  e.g = g; e.bd = bit_depth; e.qidx = base_q_idx; e.rnd = quant_rnd;
  e.src[0] = src_y; e.src[1] = src_u; e.src[2] = src_v;
  e.sstride[0] = sy_stride; e.sstride[1] = e.sstride[2] = suv_stride;
  e.rec[0] = rec_y; e.rec[1] = rec_u; e.rec[2] = rec_v;
  e.blocks = blocks;
  e.coef[0] = coef_y; e.coef[1] = coef_u; e.coef[2] = coef_v;
  for (int tr = 0; tr < g->tile_rows; tr++)
    for (int tc = 0; tc < g->tile_cols; tc++) {
      e.mi_row_start = g->tile_row_start_sb[tr] * 16;
      e.mi_row_end = std::min(g->tile_row_start_sb[tr + 1] * 16, g->mi_rows);
      e.mi_col_start = g->tile_col_start_sb[tc] * 16;
      e.mi_col_end = std::min(g->tile_col_start_sb[tc + 1] * 16, g->mi_cols);
      for (int r = e.mi_row_start; r < e.mi_row_end; r += 16)
        for (int c = e.mi_col_start; c < e.mi_col_end; c += 16) {
          e.clear_block_decoded(r, c);
          e.partition(r, c, 6, r, c, part_map);
        }
    }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Inter prediction (spec 7.11.3.3 / 7.11.3.4), single reference, no scaling, EIGHTTAP (regular)
// interpolation; blocks of width (height) <= 4 use the 4-tap kernels (spec: interpFilter index 4).
// ref: plane of the reference frame (stride rstride, picture size ref_w x ref_h in samples of this
// plane); (x, y): position of the block in the plane; mv in 1/8 LUMA samples (row, col); ss = 1 for
// the 4:2:0 chroma planes.
// ------------------------------------------------------------------------------------------------
extern "C" void orc_inter_predict(const uint16_t* ref, int rstride, int ref_w, int ref_h, int x, int y, int w, int h,
                                  int mv_row, int mv_col, int ss, int bd, uint16_t* dst, int dstride) {
  const int x16 = (x << 4) + ((2 * mv_col) >> ss), y16 = (y << 4) + ((2 * mv_row) >> ss);   // 1/16 sample units
  const int ix = x16 >> 4, iy = y16 >> 4, fx = x16 & 15, fy = y16 & 15;
  const int16_t* kx = w <= 4 ? av1t_sub_pel_filters_4[fx] : av1t_sub_pel_filters_8[fx];
  const int16_t* ky = h <= 4 ? av1t_sub_pel_filters_4[fy] : av1t_sub_pel_filters_8[fy];
  std::vector<int32_t> inter((size_t)(h + 7) * w);
  for (int r = 0; r < h + 7; r++) {
    const int yy = clampi(iy + r - 3, 0, ref_h - 1);
    for (int c = 0; c < w; c++) {
      int s = 0;
      for (int t = 0; t < 8; t++) s += kx[t] * (int)ref[(size_t)yy * rstride + clampi(ix + c + t - 3, 0, ref_w - 1)];
      inter[(size_t)r * w + c] = (s + 4) >> 3;          // Round2(sum, InterRound0 = 3)
    }
  }
  const int maxv = (1 << bd) - 1;
  for (int r = 0; r < h; r++)
    for (int c = 0; c < w; c++) {
      int s = 0;
      for (int t = 0; t < 8; t++) s += ky[t] * inter[(size_t)(r + t) * w + c];
      dst[(size_t)r * dstride + c] = (uint16_t)clampi((s + 1024) >> 11, 0, maxv);   // Round2(sum, InterRound1 = 11)
    }
}

// ------------------------------------------------------------------------------------------------
// Hierarchical motion estimation (encoder side, ours): open loop on SOURCE pictures.
//   pyramid: L1 = 2x2 box average (a+b+c+d+2)>>2 of L0 luma, L2 likewise from L1;
//   L2: every 8x8 block (32x32 luma) full search +-kR2, cost = SAD + lam2 (|dx| + |dy|);
//   L1: every 8x8 block (16x16 luma) +-2 around twice the parent vector, cost = SAD + lam1 (|dx| + |dy|);
//   L0: every 16x16 block +-2 around twice the L1 vector, cost = SAD + lambda (|dx| + |dy|), then a quarter-sample offset per axis
//       from the parabola through the SADs next to the winner.
// Candidates are visited centre first, then in raster order; a candidate replaces the best only if
// strictly cheaper.  Samples outside a picture are edge-replicated (for both pictures).
// ------------------------------------------------------------------------------------------------
extern "C" void orc_downscale2(const uint16_t* src, int sstride, int w, int h, uint16_t* dst, int dstride) {
  for (int y = 0; y < h / 2; y++)
    for (int x = 0; x < w / 2; x++) {
      const uint16_t* p = src + (size_t)(2 * y) * sstride + 2 * x;
      dst[(size_t)y * dstride + x] = (uint16_t)((p[0] + p[1] + p[sstride] + p[sstride + 1] + 2) >> 2);
    }
}

static inline int px_clamped(const uint16_t* img, int stride, int w, int h, int x, int y) {
  return img[(size_t)clampi(y, 0, h - 1) * stride + clampi(x, 0, w - 1)];
}
static int sad_block(const uint16_t* cur, const uint16_t* ref, int stride, int w, int h, int bx, int by, int n, int dx,
                     int dy) {
  int s = 0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++)
      s += abs(px_clamped(cur, stride, w, h, bx + j, by + i) - px_clamped(ref, stride, w, h, bx + j + dx, by + i + dy));
  return s;
}
// the block SAD of the motion search, exported so that tests can pin it against libaom's aom_highbd_sad*_c
extern "C" int orc_sad_block(const uint16_t* cur, const uint16_t* ref, int stride, int w, int h, int bx, int by, int n, int dx, int dy) {
  return sad_block(cur, ref, stride, w, h, bx, by, n, dx, dy);
}
enum { kR2 = 12 };
// offset of the minimum in quarter samples (-2..2) from the costs at -1, 0, +1 (0 when the three points are not convex)
static inline int subpel_parabola(int sm, int s0, int sp, int lambda) {
  const int num = sm - sp, den = 2 * (sm - 2 * s0 + sp);
  if (den <= 0) return 0;
  if ((int64_t)num * num <= (int64_t)4 * den * lambda) return 0;   // predicted gain num^2 / (4 den) must exceed lambda
  const int a = 8 * num + den, b = 2 * den;              // round(4 * num / den) = floor((8 num + den) / (2 den))
  const int q = a >= 0 ? a / b : -((-a + b - 1) / b);
  return clampi(q, -2, 2);
}
// cur/ref: three luma levels each (L0 stride = g->stride[0]; L1, L2 have strides stride[0]/2, stride[0]/4).
// mv_out: [h8*w8][2] (row, col) in 1/8 luma samples, the vector of a 16x16 block replicated on its 8x8 units.
// lambda: cost of one integer sample of deviation from the projected parent vector at L0, in SAD units
// (L1 uses lambda/4, L2 max(1, (lambda >> shift2)/16) per quarter-resolution sample from the zero vector); a
// quarter-sample offset is kept on an axis only if the parabola predicts a SAD gain above lambda.
// shift2 (bit depth - 8): the quarter-resolution search compares 8-bit samples, min(v >> shift2, 255), with lambda
// scaled alike.
extern "C" void orc_hme(const Av1bGeom* g, const uint16_t* cur0, const uint16_t* cur1, const uint16_t* cur2_in,
                        const uint16_t* ref0, const uint16_t* ref1, const uint16_t* ref2_in, int lambda, int shift2, int16_t* mv_out) {
  const int lam1 = lambda >> 2, lam2 = std::max(1, (lambda >> shift2) >> 4);
  const int W = g->width, H = g->height, s0 = g->stride[0], s1 = s0 / 2, s2 = s0 / 4;
  const int w1 = W / 2, h1 = H / 2, w2 = W / 4, h2 = H / 4;
  std::vector<uint16_t> c2((size_t)s2 * h2), r2((size_t)s2 * h2);
  for (size_t i = 0; i < c2.size(); i++) { c2[i] = (uint16_t)std::min(cur2_in[i] >> shift2, 255); r2[i] = (uint16_t)std::min(ref2_in[i] >> shift2, 255); }
  const uint16_t* cur2 = c2.data();
  const uint16_t* ref2 = r2.data();
  const int n2x = (W + 31) / 32, n2y = (H + 31) / 32, n1x = (W + 15) / 16, n1y = (H + 15) / 16;
  std::vector<int> mv2((size_t)n2x * n2y * 2), mv1((size_t)n1x * n1y * 2);
  for (int by = 0; by < n2y; by++)
    for (int bx = 0; bx < n2x; bx++) {
      int best = sad_block(cur2, ref2, s2, w2, h2, bx * 8, by * 8, 8, 0, 0), bdx = 0, bdy = 0;
      for (int dy = -kR2; dy <= kR2; dy++)
        for (int dx = -kR2; dx <= kR2; dx++) {
          if (!dx && !dy) continue;
          const int c = sad_block(cur2, ref2, s2, w2, h2, bx * 8, by * 8, 8, dx, dy) + lam2 * (abs(dx) + abs(dy));
          if (c < best) { best = c; bdx = dx; bdy = dy; }
        }
      mv2[(by * n2x + bx) * 2] = bdy; mv2[(by * n2x + bx) * 2 + 1] = bdx;
    }
  for (int by = 0; by < n1y; by++)
    for (int bx = 0; bx < n1x; bx++) {
      const int py = 2 * mv2[((by >> 1) * n2x + (bx >> 1)) * 2], pxv = 2 * mv2[((by >> 1) * n2x + (bx >> 1)) * 2 + 1];
      int best = sad_block(cur1, ref1, s1, w1, h1, bx * 8, by * 8, 8, pxv, py), bdx = 0, bdy = 0;
      for (int dy = -2; dy <= 2; dy++)
        for (int dx = -2; dx <= 2; dx++) {
          if (!dx && !dy) continue;
          const int c = sad_block(cur1, ref1, s1, w1, h1, bx * 8, by * 8, 8, pxv + dx, py + dy) + lam1 * (abs(dx) + abs(dy));
          if (c < best) { best = c; bdx = dx; bdy = dy; }
        }
      mv1[(by * n1x + bx) * 2] = py + bdy; mv1[(by * n1x + bx) * 2 + 1] = pxv + bdx;
    }
  for (int by = 0; by < n1y; by++)
    for (int bx = 0; bx < n1x; bx++) {
      const int py = 2 * mv1[(by * n1x + bx) * 2], pxv = 2 * mv1[(by * n1x + bx) * 2 + 1];
      int sads[25];
      for (int dy = -2; dy <= 2; dy++)
        for (int dx = -2; dx <= 2; dx++)
          sads[(dy + 2) * 5 + dx + 2] = sad_block(cur0, ref0, s0, W, H, bx * 16, by * 16, 16, pxv + dx, py + dy);
      int k = 12, best = sads[12];
      for (int c = 0; c < 25; c++) {
        const int cost = sads[c] + lambda * (abs(c / 5 - 2) + abs(c % 5 - 2));
        if (c != 12 && cost < best) { best = cost; k = c; }
      }
      const int bdy = k / 5 - 2, bdx = k % 5 - 2;
      // quarter-sample refinement: vertex of the parabola through the three SADs around the winner, per axis
      int qx = 0, qy = 0;
      if (k % 5 >= 1 && k % 5 <= 3) qx = subpel_parabola(sads[k - 1], sads[k], sads[k + 1], lambda);
      if (k / 5 >= 1 && k / 5 <= 3) qy = subpel_parabola(sads[k - 5], sads[k], sads[k + 5], lambda);
      const int mvy = (py + bdy) * 8 + 2 * qy, mvx = (pxv + bdx) * 8 + 2 * qx;
      for (int uy = by * 2; uy < std::min(by * 2 + 2, g->h8); uy++)
        for (int ux = bx * 2; ux < std::min(bx * 2 + 2, g->w8); ux++) {
          mv_out[(uy * g->w8 + ux) * 2] = (int16_t)mvy;
          mv_out[(uy * g->w8 + ux) * 2 + 1] = (int16_t)mvx;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Vector-field regularisation (encoder side, ours): SAD of a 16x16 block at quarter-sample vectors on a bilinear
// interpolation of the reference SOURCE picture (sad_block_q), the frame's dominant vector (orc_mv_dominant) and the
// superblock-level rate-distortion sweeps (orc_me_sbrd).
// ------------------------------------------------------------------------------------------------
static int sad_block_q(const uint16_t* cur, const uint16_t* ref, int stride, int w, int h, int bx, int by, int n, int mvx,
                       int mvy) {   // mv in 1/8 luma samples (multiples of 2)
  const int ix = mvx >> 3, iy = mvy >> 3, fx = (mvx & 7) >> 1, fy = (mvy & 7) >> 1;   // floor, quarter-sample phase 0..3
  int s = 0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      const int x = bx + j + ix, y = by + i + iy;
      int p;
      if ((fx | fy) == 0) {
        p = px_clamped(ref, stride, w, h, x, y);
      } else {
        const int a = px_clamped(ref, stride, w, h, x, y), b = px_clamped(ref, stride, w, h, x + 1, y);
        const int c = px_clamped(ref, stride, w, h, x, y + 1), d = px_clamped(ref, stride, w, h, x + 1, y + 1);
        p = ((4 - fx) * (4 - fy) * a + fx * (4 - fy) * b + (4 - fx) * fy * c + fx * fy * d + 8) >> 4;
      }
      s += abs(px_clamped(cur, stride, w, h, bx + j, by + i) - p);
    }
  return s;
}

// Dominant vector of a field of n vectors: 1024-bin histogram over a hash of the vector; the bin with the most entries
// wins (lowest bin on ties) and stands for the largest packed vector that fell into it.
static inline uint32_t mv_pack(int mvy, int mvx) { return ((uint32_t)(uint16_t)(int16_t)mvy << 16) | (uint16_t)(int16_t)mvx; }
static inline uint32_t mv_hash(uint32_t k) { return ((k * 2654435761u) >> 22) & 1023u; }
extern "C" void orc_mv_dominant(const int16_t* mv, int n, int16_t* dom) {
  std::vector<uint32_t> cnt(1024, 0), key(1024, 0);
  for (int i = 0; i < n; i++) {
    const uint32_t k = mv_pack(mv[2 * i], mv[2 * i + 1]), b = mv_hash(k);
    cnt[b]++;
    key[b] = std::max(key[b], k ^ 0x80008000u);   // order-preserving for signed halves
  }
  int best = 0;
  for (int b = 1; b < 1024; b++) if (cnt[b] > cnt[best]) best = b;
  const uint32_t k = key[best] ^ 0x80008000u;
  dom[0] = (int16_t)(k >> 16); dom[1] = (int16_t)(k & 0xFFFF);
}

// Superblock-level rate-distortion regularisation of the vector field (orc_me_sbrd).
// mv_io: [h8*w8][2] as produced by orc_hme (the vector of a 16x16 block replicated on its 8x8 units); updated in place.
//
// The field of 16x16-block vectors is swept `passes` times; a sweep visits the 64x64 superblocks of one checkerboard
// colour, then those of the other (superblocks of one colour share no edge, so within a half sweep they are independent
// of each other and may run in any order or in parallel; every half sweep sees the result of the one before).
// For a superblock (up to 4x4 blocks):
//   candidates, in this order and without repeats: zero, the frame's dominant vector (of the field the first sweep
//     starts from), the vector of the block left of / above / right of / below the superblock's first block row / column
//     (where the picture has one), the superblock's own vectors in raster order;
//   T[c][b] = SAD of block b at candidate c (sad_block_q);
//   relaxation, blocks in raster order: b takes the candidate with the smallest T[c][b] + lam_s * (number of its four
//     neighbours inside the picture that hold another vector), first candidate on ties;
//   merge at 64x64: the cost of the field as it stands, sum over the blocks of T[own][b] + lam_r * bits(b) with
//     bits = 1 when the block left of or above b holds the same vector, else 12, against the best single candidate for
//     the whole superblock, sum of T[c][b] + lam_r * (1 when c is zero or held by the block left of / above the
//     superblock's first block, else 12): the single candidate is taken when it is strictly cheaper;
//   else the same test for each of the four 32x32 quadrants in raster order.
// Flat / noisy areas, where the SADs of all candidates are within noise of each other, collapse onto one vector per
// superblock (and, through the neighbour candidates, per region): skipped blocks merge to 32x32 / 64x64, and the vectors
// that are coded are predicted exactly (NEARESTMV).
extern "C" void orc_me_sbrd(const Av1bGeom* g, const uint16_t* cur0, const uint16_t* ref0, int lam_s, int lam_r, int passes,
                            int16_t* mv_io) {
  const int W = g->width, H = g->height, s0 = g->stride[0];
  const int n1x = (W + 15) / 16, n1y = (H + 15) / 16, n1 = n1x * n1y;
  std::vector<int16_t> v((size_t)n1 * 2);
  for (int by = 0; by < n1y; by++)
    for (int bx = 0; bx < n1x; bx++) {
      const int16_t* m = mv_io + ((size_t)(by * 2) * g->w8 + bx * 2) * 2;
      v[(by * n1x + bx) * 2] = m[0]; v[(by * n1x + bx) * 2 + 1] = m[1];
    }
  int16_t dom[2];
  orc_mv_dominant(v.data(), n1, dom);
  auto pk = [&](int by, int bx) { return mv_pack(v[(by * n1x + bx) * 2], v[(by * n1x + bx) * 2 + 1]); };
  const int nsx = (n1x + 3) / 4, nsy = (n1y + 3) / 4;
  for (int half = 0; half < 2 * passes; half++)
    for (int sby = 0; sby < nsy; sby++)
      for (int sbx = 0; sbx < nsx; sbx++) {
        if (((sbx + sby) & 1) != (half & 1)) continue;
        const int y0 = sby * 4, x0 = sbx * 4, nr = std::min(4, n1y - y0), nc = std::min(4, n1x - x0);
        uint32_t cand[22]; int ncand = 0;
        auto add = [&](uint32_t k) { for (int i = 0; i < ncand; i++) if (cand[i] == k) return; cand[ncand++] = k; };
        add(mv_pack(0, 0)); add(mv_pack(dom[0], dom[1]));
        if (x0 > 0) add(pk(y0, x0 - 1));
        if (y0 > 0) add(pk(y0 - 1, x0));
        if (x0 + nc < n1x) add(pk(y0, x0 + nc));
        if (y0 + nr < n1y) add(pk(y0 + nr, x0));
        for (int r = 0; r < nr; r++) for (int c = 0; c < nc; c++) add(pk(y0 + r, x0 + c));
        int T[22][16];
        for (int k = 0; k < ncand; k++)
          for (int r = 0; r < nr; r++)
            for (int c = 0; c < nc; c++)
              T[k][r * 4 + c] = sad_block_q(cur0, ref0, s0, W, H, (x0 + c) * 16, (y0 + r) * 16, 16, (int16_t)(cand[k] & 0xFFFF), (int16_t)(cand[k] >> 16));
        auto set = [&](int by, int bx, uint32_t k) { v[(by * n1x + bx) * 2] = (int16_t)(k >> 16); v[(by * n1x + bx) * 2 + 1] = (int16_t)(k & 0xFFFF); };
        auto idx_of = [&](uint32_t k) { for (int i = 0; i < ncand; i++) if (cand[i] == k) return i; return 0; };
        // relaxation
        for (int r = 0; r < nr; r++)
          for (int c = 0; c < nc; c++) {
            const int by = y0 + r, bx = x0 + c;
            int best = 0, bk = -1;
            for (int k = 0; k < ncand; k++) {
              int diff = 0;
              if (bx > 0) diff += pk(by, bx - 1) != cand[k];
              if (bx + 1 < n1x) diff += pk(by, bx + 1) != cand[k];
              if (by > 0) diff += pk(by - 1, bx) != cand[k];
              if (by + 1 < n1y) diff += pk(by + 1, bx) != cand[k];
              const int cost = T[k][r * 4 + c] + lam_s * diff;
              if (bk < 0 || cost < best) { best = cost; bk = k; }
            }
            set(by, bx, cand[bk]);
          }
        // merge tests
        auto test = [&](int r0, int c0, int r1, int c1) {
          int jc = 0;
          for (int r = r0; r < r1; r++)
            for (int c = c0; c < c1; c++) {
              const int by = y0 + r, bx = x0 + c;
              const uint32_t k = pk(by, bx);
              const bool same = (bx > 0 && pk(by, bx - 1) == k) || (by > 0 && pk(by - 1, bx) == k);
              jc += T[idx_of(k)][r * 4 + c] + lam_r * (same ? 1 : 12);
            }
          int best = 0, bk = -1;
          const int by = y0 + r0, bx = x0 + c0;
          for (int k = 0; k < ncand; k++) {
            const bool nb = cand[k] == mv_pack(0, 0) || (bx > 0 && pk(by, bx - 1) == cand[k]) || (by > 0 && pk(by - 1, bx) == cand[k]);
            int j = lam_r * (nb ? 1 : 12);
            for (int r = r0; r < r1; r++) for (int c = c0; c < c1; c++) j += T[k][r * 4 + c];
            if (bk < 0 || j < best) { best = j; bk = k; }
          }
          if (best < jc) {
            for (int r = r0; r < r1; r++) for (int c = c0; c < c1; c++) set(y0 + r, x0 + c, cand[bk]);
            return true;
          }
          return false;
        };
        if (!test(0, 0, nr, nc))
          for (int qr = 0; qr < nr; qr += 2)
            for (int qc = 0; qc < nc; qc += 2) test(qr, qc, std::min(qr + 2, nr), std::min(qc + 2, nc));
      }
  for (int by = 0; by < n1y; by++)
    for (int bx = 0; bx < n1x; bx++)
      for (int uy = by * 2; uy < std::min(by * 2 + 2, g->h8); uy++)
        for (int ux = bx * 2; ux < std::min(bx * 2 + 2, g->w8); ux++) {
          mv_io[(uy * g->w8 + ux) * 2] = v[(by * n1x + bx) * 2];
          mv_io[(uy * g->w8 + ux) * 2 + 1] = v[(by * n1x + bx) * 2 + 1];
        }
}

// ------------------------------------------------------------------------------------------------
// Scene-change score between two source pictures (encoder side, ours): sum of absolute luma differences on the 1/8 x 1/8
// sample grid that starts at (4, 4).  The decision rule on top of it is integer only (chain.py scene_cuts / encoder.cc):
// a cut = a score above 10 per sample (8-bit units) AND above three times the running level of change plus 2 per sample,
// at least 12 frames after the last key frame; the running level is (4 * level + score) / 5, restarted at a cut.
// ------------------------------------------------------------------------------------------------
extern "C" uint32_t orc_scene_score(const Av1bGeom* g, const uint16_t* cur_y, const uint16_t* prev_y, int stride) {
  uint32_t s = 0;
  for (int y = 4; y < g->height; y += 8)
    for (int x = 4; x < g->width; x += 8) s += (uint32_t)abs((int)cur_y[(size_t)y * stride + x] - (int)prev_y[(size_t)y * stride + x]);
  return s;
}

// ------------------------------------------------------------------------------------------------
// Noise level of a source picture (encoder side, ours; Immerkaer's fast estimate made robust by a percentile): for
// every 16x16 luma block that lies inside the picture, B = sum over its 14x14 inner samples of |I * N| with the
// noise-sensitive mask N = [1 -2 1; -2 4 -2; 1 -2 1] (it cancels constant, linear and many quadratic structures).  The
// blocks go into a histogram of 4096 bins of width 16; the result is the centre of the bin that holds the lower
// quartile of the blocks outside bin 0 (saturated areas and letterbox bars carry no noise at all and say nothing about
// the rest of the picture); 0 when every block is in bin 0.  sigma ~= sqrt(pi / 2) * B / (6 * 196).
// ------------------------------------------------------------------------------------------------
extern "C" int orc_noise_estimate(const Av1bGeom* g, const uint16_t* src_y, int stride) {
  std::vector<uint32_t> hist(4096, 0);
  uint32_t n = 0;
  for (int by = 0; by * 16 + 16 <= g->height; by++)
    for (int bx = 0; bx * 16 + 16 <= g->width; bx++) {
      int64_t B = 0;
      for (int i = 1; i < 15; i++)
        for (int j = 1; j < 15; j++) {
          const uint16_t* p = src_y + (size_t)(by * 16 + i) * stride + bx * 16 + j;
          const int v = (int)p[-stride - 1] - 2 * (int)p[-stride] + (int)p[-stride + 1] - 2 * (int)p[-1] + 4 * (int)p[0] - 2 * (int)p[1] +
                        (int)p[stride - 1] - 2 * (int)p[stride] + (int)p[stride + 1];
          B += abs(v);
        }
      hist[std::min<int64_t>(B >> 4, 4095)]++;
      n++;
    }
  n -= hist[0];
  if (n == 0) return 0;
  const uint32_t want = (n + 3) / 4;
  uint32_t acc = 0;
  for (int b = 1; b < 4096; b++) {
    acc += hist[b];
    if (acc >= want) return (b << 4) + 8;
  }
  return (4095 << 4) + 8;
}

// ------------------------------------------------------------------------------------------------
// Motion-compensated temporal filter of a key / anchor SOURCE picture (encoder side, ours): the picture is replaced
// by a weighted mean of itself (weight 256) and of its neighbours in time, each motion-compensated onto it with the
// normative interpolation (16x16 luma blocks, vectors mvs[k] from the motion search of the picture against neighbour
// k).  Per neighbour and block:  mse = SSE_luma / samples;  wb = clamp(16 - 16 mse / thr_b, 0, 16);  per sample of
// every plane  w = wb * clamp(16 - 16 d^2 / thr_p, 0, 16)  with d = prediction - picture.  Noise that is independent
// from frame to frame averages out, anything the vectors do not explain keeps the picture's own samples.
// ------------------------------------------------------------------------------------------------
extern "C" void orc_mctf(const Av1bGeom* g, int bd, const uint16_t* cur_y, const uint16_t* cur_u, const uint16_t* cur_v,
                         int n_nb, const uint16_t* const* nb_planes /*[n_nb*3]*/, const int16_t* const* mvs /*[n_nb]*/,
                         int thr_b, int thr_p, uint16_t* out_y, uint16_t* out_u, uint16_t* out_v) {
  const uint16_t* cur[3] = {cur_y, cur_u, cur_v};
  uint16_t* out[3] = {out_y, out_u, out_v};
  const int W = g->width, H = g->height;
  std::vector<uint32_t> num[3], den[3];
  for (int p = 0; p < 3; p++) {
    const size_t n = (size_t)g->stride[p] * g->rows[p];
    num[p].assign(n, 0); den[p].assign(n, 256);
    for (size_t i = 0; i < n; i++) num[p][i] = 256u * cur[p][i];
  }
  uint16_t pred[3][16 * 16];
  for (int k = 0; k < n_nb; k++)
    for (int by = 0; by * 16 < H; by++)
      for (int bx = 0; bx * 16 < W; bx++) {
        const int16_t* mv = mvs[k] + ((size_t)(by * 2) * g->w8 + bx * 2) * 2;
        for (int p = 0; p < 3; p++) {
          const int ss = p > 0, n = 16 >> ss;
          orc_inter_predict(nb_planes[k * 3 + p], g->stride[p], W >> ss, H >> ss, (bx * 16) >> ss, (by * 16) >> ss, n, n, mv[0],
                            mv[1], ss, bd, pred[p], n);
        }
        const int bw = std::min(16, W - bx * 16), bh = std::min(16, H - by * 16);
        int64_t sse = 0;
        for (int i = 0; i < bh; i++)
          for (int j = 0; j < bw; j++) {
            const int d = (int)pred[0][i * 16 + j] - (int)cur[0][(size_t)(by * 16 + i) * g->stride[0] + bx * 16 + j];
            sse += d * d;
          }
        const int mse = (int)(sse / (bw * bh));
        const int wb = clampi(16 - (int)(((int64_t)16 * mse) / thr_b), 0, 16);
        if (wb == 0) continue;
        for (int p = 0; p < 3; p++) {
          const int ss = p > 0, n = 16 >> ss;
          for (int i = 0; i < (bh >> ss); i++)
            for (int j = 0; j < (bw >> ss); j++) {
              const size_t o = (size_t)(((by * 16) >> ss) + i) * g->stride[p] + ((bx * 16) >> ss) + j;
              const int d = (int)pred[p][i * n + j] - (int)cur[p][o];
              const int w = wb * clampi(16 - (16 * d * d) / thr_p, 0, 16);
              num[p][o] += (uint32_t)w * pred[p][i * n + j];
              den[p][o] += (uint32_t)w;
            }
        }
      }
  for (int p = 0; p < 3; p++) {
    const size_t n = (size_t)g->stride[p] * g->rows[p];
    for (size_t i = 0; i < n; i++) out[p][i] = (uint16_t)((num[p][i] + den[p][i] / 2) / den[p][i]);
  }
}

// ------------------------------------------------------------------------------------------------
// Inter frame encode: every block is predicted from the previous reconstructed frame with the given
// vector (one per 8x8 unit; all units of a block carry the same vector), DCT_DCT residual coding with
// the same quantiser as the intra path.  Blocks are independent of each other.
// ------------------------------------------------------------------------------------------------
// tb_zero_thr: a transform block whose quantised levels sum (in magnitude) to at most the threshold is
// dropped (all levels zero): thr for 16x16 blocks, thr/2 for 8x8, never for 4x4.
extern "C" int orc_encode_inter_frame(const Av1bGeom* g, int bit_depth, int base_q_idx, int quant_rnd, int tb_zero_thr,
                                      const uint16_t* src_y, const uint16_t* src_u, const uint16_t* src_v,
                                      int sy_stride, int suv_stride, const uint8_t* part_map, const int16_t* mvs,
                                      const uint16_t* ref_y, const uint16_t* ref_u, const uint16_t* ref_v,
                                      uint16_t* rec_y, uint16_t* rec_u, uint16_t* rec_v, Av1bBlockInfo* blocks,
                                      int16_t* coef_y, int16_t* coef_u, int16_t* coef_v) {
  const uint16_t* src[3] = {src_y, src_u, src_v};
  const int sstride[3] = {sy_stride, suv_stride, suv_stride};
  const uint16_t* ref[3] = {ref_y, ref_u, ref_v};
  uint16_t* rec[3] = {rec_y, rec_u, rec_v};
  int16_t* coef[3] = {coef_y, coef_u, coef_v};
  std::vector<uint16_t> pred(64 * 64);
  std::vector<int16_t> resid(64 * 64);
  std::vector<int32_t> cf(32 * 32), dq(32 * 32);
  std::vector<int16_t> lv(32 * 32);
  for (int uy = 0; uy < g->h8; uy++)
    for (int ux = 0; ux < g->w8; ux++) {
      const int bl = part_map[uy * g->w8 + ux], n8 = 1 << (bl - 3);
      if ((ux | uy) & (n8 - 1)) continue;   // not the top-left unit of its block
      Av1bBlockInfo info;
      memset(&info, 0, sizeof(info));
      info.blk_log2 = (uint8_t)bl;
      info.is_inter = 1;
      info.mv[0] = mvs[(uy * g->w8 + ux) * 2];
      info.mv[1] = mvs[(uy * g->w8 + ux) * 2 + 1];
      info.tx_type_y = AV1B_DCT_DCT;
      for (int p = 0; p < 3; p++) {
        const int ss = p > 0;
        const int n = std::min(1 << (bl - ss), ss ? 32 : 64), bn = 1 << (bl - ss);   // transform / prediction size
        const int x = (ux * 8) >> ss, y = (uy * 8) >> ss;
        const int pw = g->width >> ss, ph = g->height >> ss;
        (void)bn;
        orc_inter_predict(ref[p], g->stride[p], pw, ph, x, y, n, n, info.mv[0], info.mv[1], ss, bit_depth, pred.data(), n);
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++)
          resid[i * n + j] = (int16_t)((int)src[p][(size_t)(y + i) * sstride[p] + x + j] - (int)pred[i * n + j]);
        const int cn = std::min(n, 32);
        // early skip (encoder decision, ours): a transform block whose residual sums (in magnitude) to less than
        // dc_q * n / 16 is not coded -- a coefficient is about 0.6 * SAD at most (all of the residual in one basis
        // function), so such a block quantises to zero anyway except in contrived cases, and the transform is saved
        int sad = 0;
        for (int i = 0; i < n * n; i++) sad += abs(resid[i]);
        const int dcq = bit_depth == 8 ? av1t_dc_q_8[base_q_idx] : av1t_dc_q_10[base_q_idx];
        if (sad < ((dcq * n) >> 4)) {
          for (int i = 0; i < cn * cn; i++) { lv[i] = 0; dq[i] = 0; }
        } else {
          orc_fwd_txfm2d(resid.data(), n, cf.data(), cn, n, n, AV1B_DCT_DCT);
          orc_quant_dequant(cf.data(), cn, lv.data(), cn, dq.data(), cn, n, n, base_q_idx, bit_depth, quant_rnd, p);
        }
        {
          const int thr = n >= 16 ? tb_zero_thr : (n == 8 ? tb_zero_thr >> 1 : 0);
          int sum = 0;
          for (int i = 0; i < cn * cn; i++) sum += abs(lv[i]);
          if (sum <= thr) for (int i = 0; i < cn * cn; i++) { lv[i] = 0; dq[i] = 0; }
        }
        const int16_t* scan = cn == 4 ? av1t_scan_default_4x4 : cn == 8 ? av1t_scan_default_8x8
                              : cn == 16 ? av1t_scan_default_16x16 : av1t_scan_default_32x32;
        int eob = 0;
        for (int k = cn * cn - 1; k >= 0; k--) if (lv[scan[k]]) { eob = k + 1; break; }
        info.eob[p] = (uint16_t)eob;
        int16_t* cdst = coef[p] + av1b_coef_offset(g->sb_cols, p, x, y);
        for (int i = 0; i < cn * cn; i++) cdst[i] = lv[i];
        uint16_t* rdst = rec[p] + (size_t)y * g->stride[p] + x;
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) rdst[(size_t)i * g->stride[p] + j] = pred[i * n + j];
        if (eob > 0) orc_inv_txfm2d_add(dq.data(), cn, rdst, g->stride[p], n, n, AV1B_DCT_DCT, bit_depth);
      }
      info.skip = (info.eob[0] == 0 && info.eob[1] == 0 && info.eob[2] == 0);
      for (int yy = 0; yy < n8; yy++) for (int xx = 0; xx < n8; xx++) blocks[(uy + yy) * g->w8 + ux + xx] = info;
    }
  return 0;
}

// Bottom-up merge of skipped inter blocks: four sibling blocks of one size that are all inter, all skip
// and carry the same vector become one block of the next size (16 -> 32 -> 64).  The prediction of the
// merged block is sample for sample that of its children, so only the side information changes
// (and with it the deblocking edges: the merge runs before the loop filters).
extern "C" void orc_merge_skip_blocks(const Av1bGeom* g, Av1bBlockInfo* blocks) {
  for (int bl = 5; bl <= 6; bl++) {
    const int n8 = 1 << (bl - 3), h8 = n8 >> 1;
    for (int y0 = 0; y0 + n8 <= g->h8; y0 += n8)
      for (int x0 = 0; x0 + n8 <= g->w8; x0 += n8) {
        const Av1bBlockInfo& c0 = blocks[y0 * g->w8 + x0];
        bool ok = true;
        for (int q = 0; q < 4 && ok; q++) {
          const Av1bBlockInfo& c = blocks[(y0 + (q >> 1) * h8) * g->w8 + x0 + (q & 1) * h8];
          ok = c.blk_log2 == bl - 1 && c.is_inter && c.skip && c.mv[0] == c0.mv[0] && c.mv[1] == c0.mv[1];
        }
        if (!ok) continue;
        for (int yy = 0; yy < n8; yy++) for (int xx = 0; xx < n8; xx++) blocks[(y0 + yy) * g->w8 + x0 + xx].blk_log2 = (uint8_t)bl;
      }
  }
}

extern "C" int orc_geom_init(Av1bGeom* g, int w, int h, int tcl, int trl) { return av1b_geom_init(g, w, h, tcl, trl); }

// ------------------------------------------------------------------------------------------------
// Deblocking loop filter (spec 7.14), restated per sample line.  All blocks are intra, so every
// transform edge is filtered; levels are uniform per plane/direction (no segment or ref deltas).
// ------------------------------------------------------------------------------------------------
// px[-8..7] around the edge: px[-1] = p0, px[0] = q0.  filter_size 4, 6 (chroma), 8 or 16.
static void lf_filter_line(uint16_t* s, int step, int filter_size, int lvl, int sharp, int bd) {
  int shift = sharp > 4 ? 2 : (sharp > 0 ? 1 : 0);
  int limit = sharp > 0 ? clampi(lvl >> shift, 1, 9 - sharp) : std::max(1, lvl >> shift);
  int blimit = 2 * (lvl + 2) + limit;
  int thresh = lvl >> 4;
  const int sb = bd - 8;
  limit <<= sb; blimit <<= sb; thresh <<= sb;
  const int one = 1 << sb;
  auto P = [&](int i) -> int { return s[-(i + 1) * step]; };   // p_i
  auto Q = [&](int i) -> int { return s[i * step]; };          // q_i
  const int p0 = P(0), p1 = P(1), q0 = Q(0), q1 = Q(1);
  bool hev = abs(p1 - p0) > thresh || abs(q1 - q0) > thresh;
  bool mask = abs(p1 - p0) <= limit && abs(q1 - q0) <= limit && abs(p0 - q0) * 2 + abs(p1 - q1) / 2 <= blimit;
  bool flat = false, flat2 = false;
  if (filter_size >= 6) {
    const int p2 = P(2), q2 = Q(2);
    mask = mask && abs(p2 - p1) <= limit && abs(q2 - q1) <= limit;
    flat = abs(p1 - p0) <= one && abs(q1 - q0) <= one && abs(p2 - p0) <= one && abs(q2 - q0) <= one;
  }
  if (filter_size >= 8) {
    const int p3 = P(3), q3 = Q(3);
    mask = mask && abs(p3 - P(2)) <= limit && abs(q3 - Q(2)) <= limit;
    flat = flat && abs(p3 - p0) <= one && abs(q3 - q0) <= one;
  }
  if (filter_size == 16) {
    flat2 = abs(P(4) - p0) <= one && abs(P(5) - p0) <= one && abs(P(6) - p0) <= one &&
            abs(Q(4) - q0) <= one && abs(Q(5) - q0) <= one && abs(Q(6) - q0) <= one;
  }
  if (!mask) return;
  if (filter_size == 4 || !flat) {
    const int lo = -(1 << (bd - 1)), hi = (1 << (bd - 1)) - 1, off = 0x80 << sb;
    const int ps1 = p1 - off, ps0 = p0 - off, qs0 = q0 - off, qs1 = q1 - off;
    int f = hev ? clampi(ps1 - qs1, lo, hi) : 0;
    f = clampi(f + 3 * (qs0 - ps0), lo, hi);
    const int f1 = clampi(f + 4, lo, hi) >> 3, f2 = clampi(f + 3, lo, hi) >> 3;
    s[0] = (uint16_t)(clampi(qs0 - f1, lo, hi) + off);
    s[-step] = (uint16_t)(clampi(ps0 + f2, lo, hi) + off);
    if (!hev) {
      const int f3 = (f1 + 1) >> 1;
      s[step] = (uint16_t)(clampi(qs1 - f3, lo, hi) + off);
      s[-2 * step] = (uint16_t)(clampi(ps1 + f3, lo, hi) + off);
    }
    return;
  }
  // wide filters (spec 7.14.6.4): window of 2n+1 taps, the centre (2*n2+1) taps doubled
  int n, log2sz, n2;
  if (filter_size == 16 && flat2) { n = 6; log2sz = 4; n2 = 1; }
  else if (filter_size == 6) { n = 2; log2sz = 3; n2 = 1; }
  else { n = 3; log2sz = 3; n2 = 0; }
  int in[16], out[16];
  for (int i = -(n + 1); i <= n; i++) in[i + 8] = s[i * step];
  for (int i = -n; i < n; i++) {
    int t = 0;
    for (int j = -n; j <= n; j++) {
      const int p = clampi(i + j, -(n + 1), n);
      t += in[p + 8] * (abs(j) <= n2 ? 2 : 1);
    }
    out[i + 8] = (t + (1 << (log2sz - 1))) >> log2sz;
  }
  for (int i = -n; i < n; i++) s[i * step] = (uint16_t)out[i + 8];
}

extern "C" void orc_lf_filter_line(uint16_t* s, int step, int filter_size, int lvl, int sharp, int bd) {
  lf_filter_line(s, step, filter_size, lvl, sharp, bd);
}

// rec: padded planes (geometry strides); blocks: per-8x8 side info.  lf_level: Y-vertical-edges,
// Y-horizontal-edges, U, V (frame header loop_filter_level[0..3]).
extern "C" void orc_deblock_frame(const Av1bGeom* g, int bd, const Av1bBlockInfo* blocks, uint16_t* rec_y,
                                  uint16_t* rec_u, uint16_t* rec_v, const int32_t* lf_level, int sharp) {
  uint16_t* rec[3] = {rec_y, rec_u, rec_v};
  for (int p = 0; p < 3; p++) {
    if (p == 0 && !lf_level[0] && !lf_level[1]) break;
    if (p > 0 && !lf_level[p + 1]) continue;
    const int ss = p > 0;
    const int rows4 = g->mi_rows >> ss, cols4 = g->mi_cols >> ss;   // plane 4x4 units
    const int stride = g->stride[p];
    for (int pass = 0; pass < 2; pass++) {
      const int lvl = p == 0 ? lf_level[pass] : lf_level[p + 1];
      if (!lvl) continue;
      for (int r4 = 0; r4 < rows4; r4++)
        for (int c4 = 0; c4 < cols4; c4++) {
          // block-info unit of this 4x4 and of the previous one across the edge
          const int ur = ss ? r4 : r4 >> 1, uc = ss ? c4 : c4 >> 1;
          const int pr4 = pass ? r4 - 1 : r4, pc4 = pass ? c4 : c4 - 1;
          if (pr4 < 0 || pc4 < 0) continue;   // picture edge
          const int pur = ss ? pr4 : pr4 >> 1, puc = ss ? pc4 : pc4 >> 1;
          const int n_cur = std::min(1 << (blocks[ur * g->w8 + uc].blk_log2 - ss), ss ? 32 : 64);
          const int n_prev = std::min(1 << (blocks[pur * g->w8 + puc].blk_log2 - ss), ss ? 32 : 64);
          const int pos = (pass ? r4 : c4) * 4;
          if (pos % n_cur) continue;          // not a transform edge
          const int fs = std::min(std::min(n_cur, n_prev), ss ? 8 : 16);
          const int filter_size = ss ? (fs == 8 ? 6 : 4) : fs;
          for (int i = 0; i < 4; i++) {
            uint16_t* s = pass ? rec[p] + (size_t)(r4 * 4) * stride + c4 * 4 + i
                               : rec[p] + (size_t)(r4 * 4 + i) * stride + c4 * 4;
            lf_filter_line(s, pass ? stride : 1, filter_size, lvl, sharp, bd);
          }
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// CDEF (spec 7.15).  Input: deblocked frame; output: separate frame (all taps read the input).
// ------------------------------------------------------------------------------------------------
static const int8_t kCdefDirections[8][2][2] = {{{-1, 1}, {-2, 2}}, {{0, 1}, {-1, 2}}, {{0, 1}, {0, 2}}, {{0, 1}, {1, 2}},
                                                {{1, 1}, {2, 2}},   {{1, 0}, {2, 1}},  {{1, 0}, {2, 0}}, {{1, 0}, {2, -1}}};
static const int kCdefDivTable[9] = {0, 840, 420, 280, 210, 168, 140, 120, 105};

// 8x8 luma block direction search (spec 7.15.2). Returns direction, *var.
extern "C" int orc_cdef_find_dir(const uint16_t* img, int stride, int32_t* var, int bd) {
  int32_t cost[8] = {0}, partial[8][15];
  memset(partial, 0, sizeof(partial));
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 8; j++) {
      const int x = (img[i * stride + j] >> (bd - 8)) - 128;
      partial[0][i + j] += x;
      partial[1][i + j / 2] += x;
      partial[2][i] += x;
      partial[3][3 + i - j / 2] += x;
      partial[4][7 + i - j] += x;
      partial[5][3 - i / 2 + j] += x;
      partial[6][j] += x;
      partial[7][i / 2 + j] += x;
    }
  for (int i = 0; i < 8; i++) {
    cost[2] += partial[2][i] * partial[2][i];
    cost[6] += partial[6][i] * partial[6][i];
  }
  cost[2] *= kCdefDivTable[8];
  cost[6] *= kCdefDivTable[8];
  for (int i = 0; i < 7; i++) {
    cost[0] += (partial[0][i] * partial[0][i] + partial[0][14 - i] * partial[0][14 - i]) * kCdefDivTable[i + 1];
    cost[4] += (partial[4][i] * partial[4][i] + partial[4][14 - i] * partial[4][14 - i]) * kCdefDivTable[i + 1];
  }
  cost[0] += partial[0][7] * partial[0][7] * kCdefDivTable[8];
  cost[4] += partial[4][7] * partial[4][7] * kCdefDivTable[8];
  for (int i = 1; i < 8; i += 2) {
    for (int j = 0; j < 5; j++) cost[i] += partial[i][3 + j] * partial[i][3 + j];
    cost[i] *= kCdefDivTable[8];
    for (int j = 0; j < 3; j++)
      cost[i] += (partial[i][j] * partial[i][j] + partial[i][10 - j] * partial[i][10 - j]) * kCdefDivTable[2 * j + 2];
  }
  int best = 0, dir = 0;
  for (int d = 0; d < 8; d++) if (cost[d] > best) { best = cost[d]; dir = d; }
  *var = (best - cost[(dir + 4) & 7]) >> 10;
  return dir;
}

static inline int floor_log2(unsigned v) { int k = -1; while (v) { k++; v >>= 1; } return k; }
static inline int cdef_constrain(int diff, int threshold, int damping) {
  if (!threshold) return 0;
  const int adj = std::max(0, damping - floor_log2((unsigned)threshold));
  const int mag = abs(diff);
  const int v = clampi(threshold - (mag >> adj), 0, mag);
  return diff < 0 ? -v : v;
}

// one block (w x h samples at (x0,y0) of a plane). in: deblocked plane; out: CDEF plane.
static void cdef_filter_block(const uint16_t* in, uint16_t* out, int stride, int x0, int y0, int w, int h, int plane_w,
                              int plane_h, int pri, int sec, int damping, int dir, int coeff_shift) {
  for (int i = 0; i < h; i++)
    for (int j = 0; j < w; j++) {
      const int x = in[(size_t)(y0 + i) * stride + x0 + j];
      int sum = 0, mx = x, mn = x;
      for (int k = 0; k < 2; k++)
        for (int sign = -1; sign <= 1; sign += 2) {
          {
            const int yy = y0 + i + sign * kCdefDirections[dir][k][0], xx = x0 + j + sign * kCdefDirections[dir][k][1];
            if (yy >= 0 && yy < plane_h && xx >= 0 && xx < plane_w) {
              const int p = in[(size_t)yy * stride + xx];
              sum += av1t_cdef_pri_taps[(pri >> coeff_shift) & 1][k] * cdef_constrain(p - x, pri, damping);
              mx = std::max(mx, p); mn = std::min(mn, p);
            }
          }
          for (int doff = -2; doff <= 2; doff += 4) {
            const int d2 = (dir + doff) & 7;
            const int yy = y0 + i + sign * kCdefDirections[d2][k][0], xx = x0 + j + sign * kCdefDirections[d2][k][1];
            if (yy >= 0 && yy < plane_h && xx >= 0 && xx < plane_w) {
              const int s = in[(size_t)yy * stride + xx];
              sum += av1t_cdef_sec_taps[k] * cdef_constrain(s - x, sec, damping);
              mx = std::max(mx, s); mn = std::min(mn, s);
            }
          }
        }
      out[(size_t)(y0 + i) * stride + x0 + j] = (uint16_t)clampi(x + ((8 + sum - (sum < 0)) >> 4), mn, mx);
    }
}

// Whole frame. in[3]: deblocked padded planes, out[3]: result (must be a different buffer).
// cdef_idx: per 64x64 strength index; the per-SB "-1" (no non-skip block) state is derived from the
// block info exactly as the decoder derives it (read_cdef is only reached by non-skip blocks).
extern "C" void orc_cdef_frame(const Av1bGeom* g, int bd, const Av1bBlockInfo* blocks, const Av1bFrameParams* fp,
                               const uint8_t* cdef_idx, const uint16_t* in_y, const uint16_t* in_u,
                               const uint16_t* in_v, uint16_t* out_y, uint16_t* out_u, uint16_t* out_v) {
  const uint16_t* in[3] = {in_y, in_u, in_v};
  uint16_t* out[3] = {out_y, out_u, out_v};
  const int cs = bd - 8;
  for (int p = 0; p < 3; p++) memcpy(out[p], in[p], (size_t)g->stride[p] * g->rows[p] * 2);
  for (int r8 = 0; r8 < g->h8; r8++)
    for (int c8 = 0; c8 < g->w8; c8++) {
      const Av1bBlockInfo& b = blocks[r8 * g->w8 + c8];
      if (b.skip) continue;
      const int idx = cdef_idx[(r8 >> 3) * g->sb_cols + (c8 >> 3)];
      int32_t var;
      const int ydir = orc_cdef_find_dir(in[0] + (size_t)(r8 * 8) * g->stride[0] + c8 * 8, g->stride[0], &var, bd);
      int pri = (fp->cdef_y_strength[idx] >> 2) << cs;
      int sec = fp->cdef_y_strength[idx] & 3; if (sec == 3) sec = 4; sec <<= cs;
      int dir = pri == 0 ? 0 : ydir;
      const int var_str = (var >> 6) ? std::min(floor_log2((unsigned)(var >> 6)), 12) : 0;
      pri = var ? (pri * (4 + var_str) + 8) >> 4 : 0;
      cdef_filter_block(in[0], out[0], g->stride[0], c8 * 8, r8 * 8, 8, 8, g->mi_cols * 4, g->mi_rows * 4, pri, sec,
                        fp->cdef_damping + cs, dir, cs);
      pri = (fp->cdef_uv_strength[idx] >> 2) << cs;
      sec = fp->cdef_uv_strength[idx] & 3; if (sec == 3) sec = 4; sec <<= cs;
      dir = pri == 0 ? 0 : ydir;
      for (int p = 1; p < 3; p++)
        cdef_filter_block(in[p], out[p], g->stride[p], c8 * 4, r8 * 4, 4, 4, g->mi_cols * 2, g->mi_rows * 2, pri, sec,
                          fp->cdef_damping + cs - 1, dir, cs);
    }
}

// Encoder-side CDEF decision (ours): per 64x64 superblock the preset (index into the frame's
// 2^cdef_bits strength pairs) with the smallest sum of squared errors against the source over the
// Y, U and V samples on the even rows of the non-skip 8x8 blocks; ties go to the lowest index.  src: padded planes.
extern "C" void orc_cdef_search(const Av1bGeom* g, int bd, const Av1bBlockInfo* blocks, const Av1bFrameParams* fp,
                                const uint16_t* in_y, const uint16_t* in_u, const uint16_t* in_v,
                                const uint16_t* src_y, const uint16_t* src_u, const uint16_t* src_v,
                                uint8_t* cdef_idx) {
  const uint16_t* src[3] = {src_y, src_u, src_v};
  const int nsb = g->sb_rows * g->sb_cols, ncand = 1 << fp->cdef_bits;
  std::vector<uint64_t> best(nsb, ~(uint64_t)0);
  std::vector<uint8_t> idx(nsb);
  std::vector<uint16_t> out[3];
  for (int p = 0; p < 3; p++) out[p].resize((size_t)g->stride[p] * g->rows[p]);
  for (int i = 0; i < nsb; i++) cdef_idx[i] = 0;
  for (int cand = 0; cand < ncand; cand++) {
    std::fill(idx.begin(), idx.end(), (uint8_t)cand);
    orc_cdef_frame(g, bd, blocks, fp, idx.data(), in_y, in_u, in_v, out[0].data(), out[1].data(), out[2].data());
    std::vector<uint64_t> sse(nsb, 0);
    for (int r8 = 0; r8 < g->h8; r8++)
      for (int c8 = 0; c8 < g->w8; c8++) {
        if (blocks[r8 * g->w8 + c8].skip) continue;
        uint64_t e = 0;
        for (int p = 0; p < 3; p++) {
          const int n = p ? 4 : 8;
          for (int i = 0; i < n; i += 2)      // the decision looks at a checkerboard of the even rows (a quarter of the samples)
            for (int j = (i >> 1) & 1; j < n; j += 2) {
              const size_t o = (size_t)(r8 * n + i) * g->stride[p] + c8 * n + j;
              const int d = (int)out[p][o] - (int)src[p][o];
              e += (uint64_t)(d * d);
            }
        }
        sse[(r8 >> 3) * g->sb_cols + (c8 >> 3)] += e;
      }
    for (int i = 0; i < nsb; i++)
      if (sse[i] < best[i]) { best[i] = sse[i]; cdef_idx[i] = (uint8_t)cand; }
  }
}

// ------------------------------------------------------------------------------------------------
// Loop restoration (spec 7.17): Wiener and self-guided filters, 64-row stripes offset by 8 luma
// rows; outside the stripe the DEBLOCKED (pre-CDEF) rows are used, at most 2 rows deep.
// ------------------------------------------------------------------------------------------------
struct LrPlane {
  const uint16_t* cdef;      // CDEF output
  const uint16_t* deblocked; // pre-CDEF
  int stride, plane_w, plane_h;
  int stripe_start, stripe_end;
  inline int sample(int x, int y) const {
    x = clampi(x, 0, plane_w - 1);
    y = clampi(y, 0, plane_h - 1);
    if (y < stripe_start) { y = std::max(stripe_start - 2, y); return deblocked[(size_t)y * stride + x]; }
    if (y > stripe_end) { y = std::min(stripe_end + 2, y); return deblocked[(size_t)y * stride + x]; }
    return cdef[(size_t)y * stride + x];
  }
};

static inline int count_units(int unit, int size) { return std::max((size + (unit >> 1)) / unit, 1); }

extern "C" void orc_lr_unit_grid(const Av1bGeom* g, const Av1bFrameParams* fp, int plane, int32_t* unit_size,
                                 int32_t* unit_rows, int32_t* unit_cols) {
  const int ss = plane > 0;
  int us = 64 << fp->lr_unit_shift;
  if (ss) us >>= fp->lr_uv_shift;
  *unit_size = us;
  *unit_rows = count_units(us, (g->height + ss) >> ss);
  *unit_cols = count_units(us, (g->width + ss) >> ss);
}

static void wiener_taps(const int8_t* c, int* f) {   // c: 3 coded taps -> 7-tap symmetric filter
  f[3] = 128;
  for (int i = 0; i < 3; i++) { f[i] = c[i]; f[6 - i] = c[i]; f[3] -= 2 * c[i]; }
}

// self-guided filter for one sample region is evaluated sample by sample through A/B planes that
// are computed per 4x4 block with a 1-sample apron (spec 7.17.3 box filter process).
static void sgr_box(const LrPlane& P, int x0, int y0, int w, int h, int r, int s, int bd, int32_t* A, int32_t* B,
                    int astride) {
  // A,B cover rows -1..h, cols -1..w  (index (i+1)*astride + (j+1))
  const int n = (2 * r + 1) * (2 * r + 1);
  const int one_by_n = av1t_one_by_x[n - 1];
  for (int i = -1; i <= h; i++)
    for (int j = -1; j <= w; j++) {
      uint32_t a = 0, b = 0;
      for (int dy = -r; dy <= r; dy++)
        for (int dx = -r; dx <= r; dx++) {
          const uint32_t c = (uint32_t)P.sample(x0 + j + dx, y0 + i + dy);
          a += c * c; b += c;
        }
      const uint32_t a_r = (bd > 8) ? (a + (1u << (2 * (bd - 8) - 1))) >> (2 * (bd - 8)) : a;
      const uint32_t d = (bd > 8) ? (b + (1u << (bd - 8 - 1))) >> (bd - 8) : b;
      const uint32_t p = (a_r * n < d * d) ? 0 : a_r * n - d * d;
      const uint32_t z = (uint32_t)(((uint64_t)p * (uint32_t)s + (1u << 19)) >> 20);
      const uint32_t a2 = av1t_x_by_xplus1[std::min<uint32_t>(z, 255)];
      const uint32_t b2 = (256 - a2) * b * (uint32_t)one_by_n;
      A[(i + 1) * astride + j + 1] = (int32_t)a2;
      B[(i + 1) * astride + j + 1] = (int32_t)((b2 + (1u << 11)) >> 12);
    }
}

// Whole frame.  cdef[3]: CDEF output; deb[3]: deblocked (pre-CDEF); out[3]: result.
// units[p]: [unit_rows][unit_cols] parameters (NULL when lr_type[p] == NONE).
extern "C" void orc_lr_frame(const Av1bGeom* g, int bd, const Av1bFrameParams* fp, const uint16_t* cdef_y,
                             const uint16_t* cdef_u, const uint16_t* cdef_v, const uint16_t* deb_y,
                             const uint16_t* deb_u, const uint16_t* deb_v, uint16_t* out_y, uint16_t* out_u,
                             uint16_t* out_v, const Av1bLrUnit* units_y, const Av1bLrUnit* units_u,
                             const Av1bLrUnit* units_v) {
  const uint16_t* cdef[3] = {cdef_y, cdef_u, cdef_v};
  const uint16_t* deb[3] = {deb_y, deb_u, deb_v};
  uint16_t* out[3] = {out_y, out_u, out_v};
  const Av1bLrUnit* units[3] = {units_y, units_u, units_v};
  const int round0 = 3, round1 = 11;   // 8/10-bit (spec 7.17.2 rounding variables)
  for (int p = 0; p < 3; p++) {
    memcpy(out[p], cdef[p], (size_t)g->stride[p] * g->rows[p] * 2);
    if (fp->lr_type[p] == AV1B_RESTORE_NONE || !units[p]) continue;
    const int ss = p > 0;
    int us, urows, ucols;
    orc_lr_unit_grid(g, fp, p, &us, &urows, &ucols);
    LrPlane P;
    P.cdef = cdef[p]; P.deblocked = deb[p]; P.stride = g->stride[p];
    P.plane_w = (g->width + ss) >> ss; P.plane_h = (g->height + ss) >> ss;
    const int bw = 4 >> 0, bh = 4 >> 0;   // process 4x4 sample blocks of the plane (MI_SIZE >> ss would be 2 for chroma;
                                           // any tiling that does not straddle a stripe or a unit gives the same result)
    for (int y = 0; y < P.plane_h; y += bh)
      for (int x = 0; x < P.plane_w; x += bw) {
        const int luma_y = y << ss;
        const int stripe = (luma_y + 8) / 64;
        P.stripe_start = (-8 + stripe * 64) >> ss;
        P.stripe_end = P.stripe_start + (64 >> ss) - 1;
        const int urow = std::min(urows - 1, ((luma_y + 8) >> ss) / us);
        const int ucol = std::min(ucols - 1, x / us);
        const Av1bLrUnit& u = units[p][urow * ucols + ucol];
        const int w = std::min(bw, P.plane_w - x), h = std::min(bh, P.plane_h - y);
        if (u.type == AV1B_RESTORE_WIENER) {
          int vf[7], hf[7];
          wiener_taps(u.wiener_v, vf);
          wiener_taps(u.wiener_h, hf);
          const int offset = 1 << (bd + 7 - round0 - 1), limit = (1 << (bd + 1 + 7 - round0)) - 1;
          int inter[10][4];
          for (int r = 0; r < h + 6; r++)
            for (int c = 0; c < w; c++) {
              int s = 0;
              for (int t = 0; t < 7; t++) s += hf[t] * P.sample(x + c + t - 3, y + r - 3);
              const int v = (s + (1 << (round0 - 1))) >> round0;
              inter[r][c] = clampi(v, -offset, limit - offset);
            }
          for (int r = 0; r < h; r++)
            for (int c = 0; c < w; c++) {
              int s = 0;
              for (int t = 0; t < 7; t++) s += vf[t] * inter[r + t][c];
              const int v = (s + (1 << (round1 - 1))) >> round1;
              out[p][(size_t)(y + r) * P.stride + x + c] = (uint16_t)clampi(v, 0, (1 << bd) - 1);
            }
        } else if (u.type == AV1B_RESTORE_SGRPROJ) {
          const int set = u.sgr_set;
          const int r0 = av1t_sgr_params[set][0], r1 = av1t_sgr_params[set][1];
          const int s0 = av1t_sgr_params[set][2], s1 = av1t_sgr_params[set][3];
          int32_t A0[6 * 6], B0[6 * 6], A1[6 * 6], B1[6 * 6];
          if (r0) sgr_box(P, x, y, w, h, r0, s0, bd, A0, B0, 6);
          if (r1) sgr_box(P, x, y, w, h, r1, s1, bd, A1, B1, 6);
          const int w0 = u.sgr_xqd[0], w1 = u.sgr_xqd[1], w2 = 128 - w0 - w1;
          for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) {
              const int src = cdef[p][(size_t)(y + i) * P.stride + x + j];
              const int uu = src << 4;
              int flt0 = uu, flt1 = uu;
              if (r0) {   // radius-2 pass: A/B only from odd rows (relative to an even origin)
                int a = 0, b = 0, shift;
                const int32_t* Ap = A0 + (i + 1) * 6 + j + 1;
                const int32_t* Bp = B0 + (i + 1) * 6 + j + 1;
                if ((y + i) & 1) {
                  shift = 4;
                  a = 6 * Ap[0] + 5 * (Ap[-1] + Ap[1]);
                  b = 6 * Bp[0] + 5 * (Bp[-1] + Bp[1]);
                } else {
                  shift = 5;
                  a = 6 * (Ap[-6] + Ap[6]) + 5 * (Ap[-7] + Ap[-5] + Ap[5] + Ap[7]);
                  b = 6 * (Bp[-6] + Bp[6]) + 5 * (Bp[-7] + Bp[-5] + Bp[5] + Bp[7]);
                }
                const int v = a * src + b;
                flt0 = (v + (1 << (8 + shift - 4 - 1))) >> (8 + shift - 4);
              }
              if (r1) {
                const int32_t* Ap = A1 + (i + 1) * 6 + j + 1;
                const int32_t* Bp = B1 + (i + 1) * 6 + j + 1;
                const int a = 4 * (Ap[0] + Ap[-1] + Ap[1] + Ap[-6] + Ap[6]) + 3 * (Ap[-7] + Ap[-5] + Ap[5] + Ap[7]);
                const int b = 4 * (Bp[0] + Bp[-1] + Bp[1] + Bp[-6] + Bp[6]) + 3 * (Bp[-7] + Bp[-5] + Bp[5] + Bp[7]);
                const int v = a * src + b;
                flt1 = (v + (1 << (8 + 5 - 4 - 1))) >> (8 + 5 - 4);
              }
              const int v = w1 * uu + w0 * flt0 + w2 * flt1;
              out[p][(size_t)(y + i) * P.stride + x + j] = (uint16_t)clampi((v + (1 << 10)) >> 11, 0, (1 << bd) - 1);
            }
        }
      }
  }
}

// ------------------------------------------------------------------------------------------------
// Encoder-side loop-restoration decision (ours; SURVEY.md 8a row E8 "search"), luma only: every restoration
// unit picks among NONE, WIENER with the taps cand->wiener_v / wiener_h and SGRPROJ with cand->sgr_set /
// sgr_xqd the one with the smallest squared error against the source over the unit's samples (a sample at
// row y belongs to unit row min(rows - 1, (y + 8) / 64), as in the filter itself); a restoring candidate must
// beat NONE by more than `bias` (the rate of its parameters in squared-error units).  Ties keep the earlier
// candidate in the order NONE, WIENER, SGRPROJ.  Frame-level type of the luma plane: SWITCHABLE.
// sse_out (may be NULL): [3][unit_rows * unit_cols] squared errors of NONE / WIENER / SGRPROJ.
// ------------------------------------------------------------------------------------------------
extern "C" void orc_lr_search(const Av1bGeom* g, int bd, const Av1bFrameParams* fp, const Av1bLrUnit* cand,
                              const uint16_t* cdef_y, const uint16_t* cdef_u, const uint16_t* cdef_v,
                              const uint16_t* deb_y, const uint16_t* deb_u, const uint16_t* deb_v,
                              const uint16_t* src_y, int64_t bias, Av1bLrUnit* units_y, uint64_t* sse_out) {
  int us, urows, ucols;
  orc_lr_unit_grid(g, fp, 0, &us, &urows, &ucols);
  const size_t n = (size_t)urows * ucols;
  std::vector<uint64_t> sse(3 * n, 0);
  std::vector<uint16_t> out[3];
  for (int p = 0; p < 3; p++) out[p].resize((size_t)g->stride[p] * g->rows[p]);
  Av1bFrameParams f1 = *fp;
  f1.lr_type[0] = AV1B_RESTORE_SWITCHABLE; f1.lr_type[1] = f1.lr_type[2] = AV1B_RESTORE_NONE;
  std::vector<Av1bLrUnit> all(n);
  for (int k = 0; k < 3; k++) {
    const uint16_t* res = cdef_y;
    if (k > 0) {
      for (size_t i = 0; i < n; i++) { all[i] = *cand; all[i].type = (int8_t)(k == 1 ? AV1B_RESTORE_WIENER : AV1B_RESTORE_SGRPROJ); }
      orc_lr_frame(g, bd, &f1, cdef_y, cdef_u, cdef_v, deb_y, deb_u, deb_v, out[0].data(), out[1].data(), out[2].data(),
                   all.data(), nullptr, nullptr);
      res = out[0].data();
    }
    for (int y = 0; y < g->height; y++) {
      const int ur = std::min(urows - 1, (y + 8) / us);
      for (int x = 0; x < g->width; x++) {
        const int uc = std::min(ucols - 1, x / us);
        const int64_t d = (int64_t)res[(size_t)y * g->stride[0] + x] - (int64_t)src_y[(size_t)y * g->stride[0] + x];
        sse[k * n + (size_t)ur * ucols + uc] += (uint64_t)(d * d);
      }
    }
  }
  for (size_t i = 0; i < n; i++) {
    int best = 0;
    uint64_t bv = sse[i];
    for (int k = 1; k < 3; k++) {
      const uint64_t v = sse[k * n + i] + (uint64_t)bias;
      if (v < bv) { bv = v; best = k; }
    }
    units_y[i] = *cand;
    units_y[i].type = (int8_t)(best == 0 ? AV1B_RESTORE_NONE : best == 1 ? AV1B_RESTORE_WIENER : AV1B_RESTORE_SGRPROJ);
  }
  if (sse_out) memcpy(sse_out, sse.data(), 3 * n * sizeof(uint64_t));
}
