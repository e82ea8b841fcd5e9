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
#include "av1_tables.h"
#include "av1_fwd_matrices.h"

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

// ------------------------------------------------------------------------------------------------
// quantiser
// ------------------------------------------------------------------------------------------------
static inline int dc_q(int qidx, int bd) { return bd == 8 ? av1t_dc_q_8[qidx] : bd == 10 ? av1t_dc_q_10[qidx] : av1t_dc_q_12[qidx]; }
static inline int ac_q(int qidx, int bd) { return bd == 8 ? av1t_ac_q_8[qidx] : bd == 10 ? av1t_ac_q_10[qidx] : av1t_ac_q_12[qidx]; }
static inline int tx_scale_shift(int w, int h) { const int p = w * h; return (p > 256) + (p > 1024); }

// Encoder quantiser (ours): level = min((|c| << s) + ((dqv * rnd) >> 7)) / dqv, 32767)
// Dequantiser (normative, spec 7.12.3): dq = ((level * dqv) & 0xFFFFFF) >> s, clipped to bd+8 bits.
// Returns eob-independent stats; levels/dq in SPEC layout.
extern "C" void orc_quant_dequant(const int32_t* coef, int cstride, int16_t* lev, int lstride, int32_t* dq,
                                  int dstride, int w, int h, int qidx, int bd, int rnd) {
  const int cw = std::min(w, 32), ch = std::min(h, 32), s = tx_scale_shift(w, h);
  const int lim = (1 << (7 + bd)) - 1;
  for (int i = 0; i < ch; i++)
    for (int j = 0; j < cw; j++) {
      const int dqv = (i | j) ? ac_q(qidx, bd) : dc_q(qidx, bd);
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

extern "C" void orc_dequant(const int16_t* lev, int lstride, int32_t* dq, int dstride, int w, int h, int qidx, int bd) {
  const int cw = std::min(w, 32), ch = std::min(h, 32), s = tx_scale_shift(w, h);
  const int lim = (1 << (7 + bd)) - 1;
  for (int i = 0; i < ch; i++)
    for (int j = 0; j < cw; j++) {
      const int dqv = (i | j) ? ac_q(qidx, bd) : dc_q(qidx, bd);
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

  // one transform block == one prediction block (TX_MODE_LARGEST, square partitions)
  void code_plane(int p, int mi_r, int mi_c, int bl, int sb_r, int sb_c, Av1bBlockInfo* bi, int mode,
                  int tx_type, bool decide, const uint8_t* cand, int ncand, int* best_mode) {
    (void)decide; (void)cand; (void)ncand; (void)best_mode; (void)bi; (void)mode; (void)tx_type;
    (void)p; (void)mi_r; (void)mi_c; (void)bl; (void)sb_r; (void)sb_c;
  }

  void edges_for(int p, int mi_r, int mi_c, int n, int sb_r, int sb_c, uint16_t* above, uint16_t* left,
                 int* have_above, int* have_left) {
    const int ss = p > 0;
    const int x = (mi_c * 4) >> ss, y = (mi_r * 4) >> ss;
    const int ha = mi_r > mi_row_start, hl = mi_c > mi_col_start;
    const int x4 = ((mi_c - sb_c) >> ss), y4 = ((mi_r - sb_r) >> ss), n4 = n >> 2;
    const int har = decoded[p][y4 - 1 + 1][x4 + n4 + 1];
    const int hbl = decoded[p][y4 + n4 + 1][x4 - 1 + 1];
    const int max_x = ((g->mi_cols * 4) >> ss) - 1, max_y = ((g->mi_rows * 4) >> ss) - 1;
    build_edges(rec[p], g->stride[p], x, y, n, n, ha, hl, har, hbl, max_x, max_y, bd, above, left);
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
      for (int p = p0; p <= p1; p++) edges_for(p, mi_r, mi_c, n, sb_r, sb_c, edge_a[p] + 1, edge_l[p] + 1, &ha, &hl);
      // mode decision: minimum SATD over the candidate list, ties -> first in list
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
        orc_quant_dequant(cf, cn, lv, cn, dq, cn, n, n, qidx, bd, rnd);
        // eob = 1 + last non-zero position in scan order
        const int16_t* scan = cn == 4 ? av1t_scan_default_4x4 : cn == 8 ? av1t_scan_default_8x8
                              : cn == 16 ? av1t_scan_default_16x16 : av1t_scan_default_32x32;
        int eob = 0;
        for (int k = cn * cn - 1; k >= 0; k--) if (lv[scan[k]]) { eob = k + 1; break; }
        info.eob[p] = (uint16_t)eob;
        int16_t* cdst = coef[p] + (size_t)y * g->stride[p] + x;
        for (int i = 0; i < cn; i++) for (int j = 0; j < cn; j++) cdst[(size_t)i * g->stride[p] + j] = lv[i * cn + j];
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

extern "C" int orc_geom_init(Av1bGeom* g, int w, int h, int tcl, int trl) { return av1b_geom_init(g, w, h, tcl, trl); }
