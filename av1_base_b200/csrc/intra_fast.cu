// Key-frame (intra) encode for the production block size (16x16 luma, 8x8 at the picture edge), split
// so that as much as possible runs frame-wide in parallel:
//
//   intra_mode_kernel   OPEN LOOP.  Intra-mode candidate evaluation for every block of the frame at once:
//                       the 13 candidate predictors are built from the SOURCE picture's neighbours and
//                       ranked by the sum of absolute 4x4-Hadamard coefficients (warp shuffles).  One CTA
//                       per superblock, one warp per block, Y and U+V decided separately.
//   intra_recon_kernel  CLOSED LOOP.  Prediction from the reconstructed neighbours with the decided mode,
//                       residual, forward transform, quantiser, normative dequantiser + inverse transform
//                       + reconstruction.  Blocks of a tile depend on each other, so one WARP walks the
//                       luma blocks of a tile and another warp its chroma blocks (the two chains never
//                       meet); inside the warp a transform block is owned by a group of N lanes exactly as
//                       in inter_kernel.cu (row t / column t per lane, warp-uniform transform operands).
//   intra_finish_kernel skip flags from the three end-of-block positions.
//
// Replaces arithmetic the reference delegates to av1an + SVT-AV1
// (/root/reference/crates/daemon/src/encode/av1an.rs:126-139).
// Must match oracle/av1_oracle.cpp orc_encode_intra_frame bit for bit (same definition as the general
// one-CTA-per-tile kernel in intra_kernel.cu, which still serves the other block sizes).
#include <cuda_runtime.h>
#include <stdint.h>
#include "av1_inv_txfm1d.h"
#include "av1_tables_dev.cuh"
#include "kernels.cuh"

namespace av1b {
using namespace av1tx;
namespace {

constexpr int kNumCand = 13;
enum { T_DCT = 0, T_ADST = 1 };

__constant__ uint8_t f_cand[kNumCand] = {AV1B_DC_PRED, AV1B_V_PRED, AV1B_H_PRED, AV1B_PAETH_PRED, AV1B_SMOOTH_PRED,
                                         AV1B_SMOOTH_V_PRED, AV1B_SMOOTH_H_PRED, AV1B_D45_PRED, AV1B_D135_PRED,
                                         AV1B_D113_PRED, AV1B_D157_PRED, AV1B_D203_PRED, AV1B_D67_PRED};
// Mode_To_Txfm restricted to what the candidates produce: (vertical type, horizontal type)
__constant__ uint8_t f_mode_vt[14] = {T_DCT, T_ADST, T_DCT, T_DCT, T_ADST, T_ADST, T_DCT, T_DCT, T_ADST, T_ADST, T_ADST, T_DCT, T_ADST, T_DCT};
__constant__ uint8_t f_mode_ht[14] = {T_DCT, T_DCT, T_ADST, T_DCT, T_ADST, T_DCT, T_ADST, T_ADST, T_DCT, T_ADST, T_DCT, T_ADST, T_ADST, T_DCT};
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ int morton(int x, int y) {
  return (x & 1) | ((y & 1) << 1) | ((x & 2) << 1) | ((y & 2) << 2) | ((x & 4) << 2) | ((y & 4) << 3) | ((x & 8) << 3) | ((y & 8) << 4);
}

struct TileCtx {
  int mi_row_start, mi_row_end, mi_col_start, mi_col_end;   // luma 4x4 units; ends clipped to the picture
};

// haveAboveRt / haveBelowLft of the block at (x4, y4) (4x4 units of the plane, relative to the superblock at
// luma mi (sb_r, sb_c)), n4 units wide: "has that position been decoded before this block" (spec 7.11.2),
// derived from the Z order instead of the sequential BlockDecoded flags the oracle keeps.
__device__ __forceinline__ void neighbour_avail(const Av1bGeom& g, const TileCtx& t, int sb_r, int sb_c, int ss, int x4,
                                                int y4, int n4, int* har, int* hbl) {
  const int U = 16 >> ss;
  const int sbw4 = (t.mi_col_end - sb_c) >> ss, sbh4 = (t.mi_row_end - sb_r) >> ss;
  const int me = morton(x4, y4);
  {
    const int ax = x4 + n4, ay = y4 - 1;
    int v;
    if (ay < 0) v = ax < sbw4;
    else if (ax >= U) v = 0;
    else v = (sb_c + (ax << ss) < g.mi_cols) && morton(ax, ay) < me;
    *har = v;
  }
  {
    const int bx = x4 - 1, by = y4 + n4;
    int v;
    if (by >= U) v = 0;
    else if (bx < 0) v = by < sbh4;
    else v = (sb_r + (by << ss) < g.mi_rows) && morton(bx, by) < me;
    *hbl = v;
  }
}

// edge preparation (spec 7.11.2 steps 1-4) by the lanes [0, nl) of a group; A = above + 1, L = left + 1
__device__ __forceinline__ void build_edges_group(const uint16_t* img, int stride, int x, int y, int n, int ha, int hl,
                                                  int har, int hbl, int max_x, int max_y, int bd, uint16_t* above,
                                                  uint16_t* left, int t, int nl) {
  const int base = 1 << (bd - 1);
  for (int i = t; i < 2 * n; i += nl) {
    uint16_t a, l;
    if (ha) {
      const int xx = (i < n || har) ? min(max_x, x + i) : min(max_x, x + n - 1);
      a = img[(size_t)(y - 1) * stride + xx];
    } else {
      a = hl ? img[(size_t)y * stride + x - 1] : (uint16_t)(base - 1);
    }
    if (hl) {
      const int yy = (i < n || hbl) ? min(max_y, y + i) : min(max_y, y + n - 1);
      l = img[(size_t)yy * stride + x - 1];
    } else {
      l = ha ? img[(size_t)(y - 1) * stride + x] : (uint16_t)(base + 1);
    }
    above[1 + i] = a;
    left[1 + i] = l;
  }
  if (t == 0) {
    uint16_t tl;
    if (ha && hl) tl = img[(size_t)(y - 1) * stride + x - 1];
    else if (ha) tl = img[(size_t)(y - 1) * stride + x];
    else if (hl) tl = img[(size_t)y * stride + x - 1];
    else tl = (uint16_t)base;
    above[0] = tl;
    left[0] = tl;
  }
}

// one predicted sample; A = above + 1, L = left + 1 (A[-1] = L[-1] = top-left)
// w: smooth weights of block size n in SHARED memory (lanes index them differently: constant memory would serialise)
__device__ __forceinline__ int pred_px(int mode, int i, int j, int n, const uint16_t* A, const uint16_t* L, int dcv,
                                       const uint8_t* w) {
  switch (mode) {
    case AV1B_DC_PRED: return dcv;
    case AV1B_V_PRED: return A[j];
    case AV1B_H_PRED: return L[i];
    case AV1B_PAETH_PRED: {
      const int tl = A[-1], base = A[j] + L[i] - tl;
      const int pl = abs(base - L[i]), pt = abs(base - A[j]), ptl = abs(base - tl);
      return (pl <= pt && pl <= ptl) ? L[i] : (pt <= ptl ? A[j] : tl);
    }
    case AV1B_SMOOTH_PRED: {
      const int s = w[i] * A[j] + (256 - w[i]) * L[n - 1] + w[j] * L[i] + (256 - w[j]) * A[n - 1];
      return (s + 256) >> 9;
    }
    case AV1B_SMOOTH_V_PRED: {
      return (w[i] * A[j] + (256 - w[i]) * L[n - 1] + 128) >> 8;
    }
    case AV1B_SMOOTH_H_PRED: {
      return (w[j] * L[i] + (256 - w[j]) * A[n - 1] + 128) >> 8;
    }
    default: break;
  }
  const int angle = tbl::mode_to_angle[mode];
  const int max_base = 2 * n - 1;
  if (angle < 90) {
    const int dx = tbl::dr_intra_derivative[angle];
    const int idx = (i + 1) * dx, base = (idx >> 6) + j, sh = (idx >> 1) & 31;
    return base < max_base ? (A[base] * (32 - sh) + A[base + 1] * sh + 16) >> 5 : A[max_base];
  } else if (angle < 180) {
    const int dx = tbl::dr_intra_derivative[180 - angle], dy = tbl::dr_intra_derivative[angle - 90];
    int idx = (j << 6) - (i + 1) * dx;
    int base = idx >> 6;
    if (base >= -1) {
      const int sh = (idx >> 1) & 31;
      return (A[base] * (32 - sh) + A[base + 1] * sh + 16) >> 5;
    }
    idx = (i << 6) - (j + 1) * dy;
    base = idx >> 6;
    const int sh = (idx >> 1) & 31;
    return (L[base] * (32 - sh) + L[base + 1] * sh + 16) >> 5;
  } else {
    const int dy = tbl::dr_intra_derivative[270 - angle];
    const int idx = (j + 1) * dy, base = (idx >> 6) + i, sh = (idx >> 1) & 31;
    return base < max_base ? (L[base] * (32 - sh) + L[base + 1] * sh + 16) >> 5 : L[max_base];
  }
}

__device__ __forceinline__ int dc_value(const uint16_t* A, const uint16_t* L, int n, int ln, int ha, int hl, int bd,
                                        int t, int nl, unsigned mask) {
  int s = 0;
  for (int i = t; i < n; i += nl) s += (ha ? A[i] : 0) + (hl ? L[i] : 0);
  for (int o = nl >> 1; o; o >>= 1) s += __shfl_xor_sync(mask, s, o);
  if (ha && hl) return (s + n) >> (ln + 1);
  if (ha || hl) return (s + (n >> 1)) >> ln;
  return 1 << (bd - 1);
}

__device__ __forceinline__ TileCtx tile_of_sb(const Av1bGeom& g, int sbx, int sby) {
  int tc = 0, tr = 0;
  while (g.tile_col_start_sb[tc + 1] <= sbx) tc++;
  while (g.tile_row_start_sb[tr + 1] <= sby) tr++;
  TileCtx t;
  t.mi_row_start = g.tile_row_start_sb[tr] * 16; t.mi_row_end = min(g.tile_row_start_sb[tr + 1] * 16, g.mi_rows);
  t.mi_col_start = g.tile_col_start_sb[tc] * 16; t.mi_col_end = min(g.tile_col_start_sb[tc + 1] * 16, g.mi_cols);
  return t;
}

// ---------------------------------------------------------------------------------------------
// open-loop mode decision
// ---------------------------------------------------------------------------------------------
struct ModeSmem {
  uint16_t above[8][2][72];    // [warp][plane slot][1 + 2n], n <= 16 luma / 8 chroma
  uint16_t left[8][2][72];
  uint8_t sw[32];              // smooth weights of sizes 4, 8, 16 at offsets 0, 4, 12 (as in the spec table)
};

// SATD cost of every candidate for nplanes planes of size n; returns the best mode (all lanes)
template <int N>
__device__ __forceinline__ int decide(const uint16_t* const src[2], int sstride, const uint16_t (*above)[72],
                                      const uint16_t (*left)[72], const int dcv[2], int nplanes, int lane,
                                      const uint8_t* sw) {
  constexpr int LN = N == 16 ? 4 : (N == 8 ? 3 : 2);
  constexpr int NPX = N * N, ITER = (NPX + 31) / 32;
  // source samples of this lane in 4x4-tile-major order (a 16-lane half warp holds one 4x4 tile)
  int sv[2][ITER];
#pragma unroll
  for (int k = 0; k < 2; k++)
#pragma unroll
    for (int it = 0; it < ITER; it++) {
      const int q = it * 32 + lane;
      sv[k][it] = 0;
      if (k < nplanes && q < NPX) {
        const int tl = q >> 4, w = q & 15, ty = tl >> (LN - 2), tx = tl & ((1 << (LN - 2)) - 1);
        sv[k][it] = src[k][(size_t)(ty * 4 + (w >> 2)) * sstride + tx * 4 + (w & 3)];
      }
    }
  int best = 0, best_cost = 0x7FFFFFFF;
  for (int m = 0; m < kNumCand; m++) {
    const int mode = f_cand[m];
    int acc = 0;
#pragma unroll
    for (int k = 0; k < 2; k++) {
      if (k >= nplanes) break;
      const uint16_t* A = above[k] + 1;
      const uint16_t* L = left[k] + 1;
#pragma unroll
      for (int it = 0; it < ITER; it++) {
        const int q = it * 32 + lane;
        int v = 0;
        if (q < NPX) {
          const int tl = q >> 4, w = q & 15, ty = tl >> (LN - 2), tx = tl & ((1 << (LN - 2)) - 1);
          v = sv[k][it] - pred_px(mode, ty * 4 + (w >> 2), tx * 4 + (w & 3), N, A, L, dcv[k], sw + N - 4);
        }
#pragma unroll
        for (int msk = 1; msk <= 8; msk <<= 1) {
          const int o = __shfl_xor_sync(0xffffffffu, v, msk);
          v = (lane & msk) ? o - v : v + o;
        }
        acc += abs(v);
      }
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (acc < best_cost) { best_cost = acc; best = mode; }
  }
  return best;
}

__global__ void __launch_bounds__(256) intra_mode_kernel(const IntraLaunch P) {
  __shared__ ModeSmem sm;
  const Av1bGeom& g = P.g;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sbx = blockIdx.x, sby = blockIdx.y, frame = blockIdx.z;
  const TileCtx T = tile_of_sb(g, sbx, sby);
  const int sb_r = sby * 16, sb_c = sbx * 16, bd = P.bit_depth;
  const uint8_t* pmap = P.part_map + (size_t)frame * P.map_elems;
  Av1bBlockInfo* blocks = P.blocks + (size_t)frame * P.map_elems;
  if (tid < 28) sm.sw[tid] = tbl::smooth_weights[tid];
  __syncthreads();
  for (int u = warp; u < 64; u += 8) {
    const int x8 = u & 7, y8 = u >> 3;
    const int mi_r = sb_r + 2 * y8, mi_c = sb_c + 2 * x8;
    if (mi_r >= g.mi_rows || mi_c >= g.mi_cols) continue;
    const int bl = pmap[(mi_r >> 1) * g.w8 + (mi_c >> 1)];
    const int n8 = 1 << (bl - 3);
    if ((x8 | y8) & (n8 - 1)) continue;             // not the top-left unit of its block
    const int ha = mi_r > T.mi_row_start, hl = mi_c > T.mi_col_start;
    int modes[2];
    for (int pass = 0; pass < 2; pass++) {
      const int ss = pass, n = 1 << (bl - ss), ln = bl - ss, nplanes = pass ? 2 : 1;
      const int x = (mi_c * 4) >> ss, y = (mi_r * 4) >> ss;
      const int max_x = ((g.mi_cols * 4) >> ss) - 1, max_y = ((g.mi_rows * 4) >> ss) - 1;
      int har, hbl;
      neighbour_avail(g, T, sb_r, sb_c, ss, (mi_c - sb_c) >> ss, (mi_r - sb_r) >> ss, n >> 2, &har, &hbl);
      const uint16_t* srcp[2];
      int dcv[2] = {0, 0};
      __syncwarp();
      for (int k = 0; k < nplanes; k++) {
        const int p = pass + k;
        const uint16_t* plane = P.src[p] + (size_t)frame * P.plane_elems[p];
        build_edges_group(plane, g.stride[p], x, y, n, ha, hl, har, hbl, max_x, max_y, bd, sm.above[warp][k], sm.left[warp][k], lane, 32);
        srcp[k] = plane + (size_t)y * g.stride[p] + x;
      }
      __syncwarp();
      for (int k = 0; k < nplanes; k++)
        dcv[k] = dc_value(sm.above[warp][k] + 1, sm.left[warp][k] + 1, n, ln, ha, hl, bd, lane, 32, 0xffffffffu);
      const int st = g.stride[pass];
      int mode;
      if (n == 16) mode = decide<16>(srcp, st, sm.above[warp], sm.left[warp], dcv, nplanes, lane, sm.sw);
      else if (n == 8) mode = decide<8>(srcp, st, sm.above[warp], sm.left[warp], dcv, nplanes, lane, sm.sw);
      else mode = decide<4>(srcp, st, sm.above[warp], sm.left[warp], dcv, nplanes, lane, sm.sw);
      modes[pass] = mode;
    }
    // side-info skeleton on every unit of the block; the closed-loop kernel adds the end-of-block positions
    Av1bBlockInfo info;
    info.blk_log2 = (uint8_t)bl; info.y_mode = (uint8_t)modes[0]; info.uv_mode = (uint8_t)modes[1]; info.skip = 0;
    info.angle_y = 0; info.angle_uv = 0; info.tx_type_y = AV1B_DCT_DCT; info.cfl_alpha_u = 0;
    info.eob[0] = info.eob[1] = info.eob[2] = 0; info.cfl_alpha_v = 0; info.is_inter = 0; info.mv[0] = info.mv[1] = 0;
    for (int o = lane; o < n8 * n8; o += 32)
      blocks[((mi_r >> 1) + o / n8) * g.w8 + (mi_c >> 1) + o % n8] = info;
  }
}

// ---------------------------------------------------------------------------------------------
// closed-loop reconstruction
// ---------------------------------------------------------------------------------------------
template <int N> struct FwdTab;
template <> struct FwdTab<4> {
  static __device__ __forceinline__ const int16_t* m(int t) { return t == T_ADST ? &tbl::fwd_adst4[0][0] : &tbl::fwd_dct4[0][0]; }
  static __device__ __forceinline__ int iscan(int p) { return tbl::iscan_default_4[p]; }
  static constexpr int kLog2 = 2, kRowShift = 0;
};
template <> struct FwdTab<8> {
  static __device__ __forceinline__ const int16_t* m(int t) { return t == T_ADST ? &tbl::fwd_adst8[0][0] : &tbl::fwd_dct8[0][0]; }
  static __device__ __forceinline__ int iscan(int p) { return tbl::iscan_default_8[p]; }
  static constexpr int kLog2 = 3, kRowShift = 1;
};
template <> struct FwdTab<16> {
  static __device__ __forceinline__ const int16_t* m(int t) { return t == T_ADST ? &tbl::fwd_adst16[0][0] : &tbl::fwd_dct16[0][0]; }
  static __device__ __forceinline__ int iscan(int p) { return tbl::iscan_default_16[p]; }
  static constexpr int kLog2 = 4, kRowShift = 2;
};

template <int N>
__device__ __forceinline__ void inv_1d(int t, int32_t* x, int range) {
  if (t == T_DCT) {
    idct<N>(x, range);
  } else {
    if constexpr (N == 4) iadst4(x, range);
    else if constexpr (N == 8) iadst8(x, range);
    else iadst16(x, range);
  }
}

struct ReconGroup {
  uint8_t sw[32];              // smooth weights (sizes 4, 8, 16 at offsets 0, 4, 12)
  uint16_t above[40];          // 1 + 2N, N <= 16
  uint16_t left[40];
  uint16_t pred[16 * 16];
  int32_t buf[16 * 17];
};

// One transform block == one prediction block, owned by lanes [0, N) + group offset of one warp.
template <int N>
__device__ __forceinline__ int intra_tb(const IntraLaunch& P, int frame, int p, int x, int y, int mode, int vt, int ht,
                                        int ha, int hl, int har, int hbl, int t, unsigned gmask, ReconGroup& G) {
  constexpr int S = N + 1, LN = FwdTab<N>::kLog2;
  const Av1bGeom& g = P.g;
  const int ss = p > 0, bd = P.bit_depth, stride = g.stride[p];
  const int max_x = ((g.mi_cols * 4) >> ss) - 1, max_y = ((g.mi_rows * 4) >> ss) - 1;
  uint16_t* recp = P.rec[p] + (size_t)frame * P.plane_elems[p];
  __syncwarp(gmask);
  build_edges_group(recp, stride, x, y, N, ha, hl, har, hbl, max_x, max_y, bd, G.above, G.left, t, N);
  __syncwarp(gmask);
  const uint16_t* A = G.above + 1;
  const uint16_t* L = G.left + 1;
  const int dcv = dc_value(A, L, N, LN, ha, hl, bd, t, N, gmask);
  // prediction + residual of row t
  {
    const uint16_t* sp = P.src[p] + (size_t)frame * P.plane_elems[p] + (size_t)(y + t) * stride + x;
#pragma unroll
    for (int c = 0; c < N; c++) {
      const int pv = pred_px(mode, t, c, N, A, L, dcv, G.sw + N - 4);
      G.pred[t * N + c] = (uint16_t)pv;
      G.buf[t * S + c] = ((int)sp[c] - pv) * 4;
    }
  }
  __syncwarp(gmask);
  // forward: column pass (lane t = column t), then row pass (lane t = row t)
  const int16_t* Fv = FwdTab<N>::m(vt);
  const int16_t* Fh = FwdTab<N>::m(ht);
  int32_t col[N];
#pragma unroll
  for (int k = 0; k < N; k++) {
    int32_t acc = 0;
#pragma unroll
    for (int i = 0; i < N; i++) acc += Fv[k * N + i] * G.buf[i * S + t];
    col[k] = (acc + 2048) >> 12;
  }
  __syncwarp(gmask);
#pragma unroll
  for (int k = 0; k < N; k++) G.buf[k * S + t] = col[k];
  __syncwarp(gmask);
  int eob = 0;
  int16_t* cdst = P.coef[p] + (size_t)frame * P.plane_elems[p] + av1b_coef_offset(g.sb_cols, p, x, y) + t * N;
  {
    int32_t row[N];
#pragma unroll
    for (int j = 0; j < N; j++) row[j] = G.buf[t * S + j];
    constexpr int sh = 24 + 2 * LN - FwdTab<N>::kRowShift - 4;
    const int lim = (1 << (7 + bd)) - 1;
    // row t of the N x N matrix of this plane's level (av1_qm_tables.h: 4x4 at 0, 8x8 at 16, 16x16 at 80) or flat
    const uint8_t* qmw = P.qm[ss] ? P.qm[ss] + (N == 4 ? 0 : N == 8 ? 16 : 80) + t * N : nullptr;
#pragma unroll 2
    for (int l = 0; l < N; l++) {
      int64_t acc = 0;
#pragma unroll
      for (int j = 0; j < N; j++) acc += (int64_t)Fh[l * N + j] * row[j];
      const int32_t c = (int32_t)((acc * 4096 + ((int64_t)1 << (sh - 1))) >> sh);
      uint32_t dqv = (t | l) ? (uint32_t)P.ac_q : (uint32_t)P.dc_q;
      if (qmw) dqv = (dqv * (uint32_t)qmw[l] + 16u) >> 5;   // quantisation matrix: Round2(q * weight, 5), spec 7.12.3
      const uint32_t a = (uint32_t)(c < 0 ? -c : c);
      uint32_t lv = (a + ((dqv * (uint32_t)P.quant_rnd) >> 7)) / dqv;
      if (lv > 32767u) lv = 32767u;
      int32_t d = (int32_t)((lv * dqv) & 0xFFFFFFu);
      if (d > lim) d = lim;
      G.buf[t * S + l] = c < 0 ? -d : d;
      cdst[l] = (int16_t)(c < 0 ? -(int32_t)lv : (int32_t)lv);
      if (lv) eob = max(eob, FwdTab<N>::iscan(t * N + l) + 1);
    }
#pragma unroll
    for (int o = N / 2; o; o >>= 1) eob = max(eob, __shfl_xor_sync(gmask, eob, o));
  }
  uint16_t* rec = recp + (size_t)y * stride + x;
  const int maxv = (1 << bd) - 1;
  if (eob == 0) {
#pragma unroll
    for (int i = 0; i < N; i++) rec[(size_t)i * stride + t] = G.pred[i * N + t];
    return 0;
  }
  const int row_range = bd + 8, col_range = max(bd + 6, 16);
  {
    int32_t dq[N];
#pragma unroll
    for (int j = 0; j < N; j++) dq[j] = sat(G.buf[t * S + j], row_range);
    inv_1d<N>(ht, dq, row_range);
#pragma unroll
    for (int j = 0; j < N; j++) {
      int32_t v = dq[j];
      if (FwdTab<N>::kRowShift > 0) v = (v + (1 << (FwdTab<N>::kRowShift > 0 ? FwdTab<N>::kRowShift - 1 : 0))) >> FwdTab<N>::kRowShift;
      G.buf[t * S + j] = v;
    }
  }
  __syncwarp(gmask);
  {
    int32_t xc[N];
#pragma unroll
    for (int i = 0; i < N; i++) xc[i] = sat(G.buf[i * S + t], col_range);
    inv_1d<N>(vt, xc, col_range);
#pragma unroll
    for (int i = 0; i < N; i++) {
      const int v = (xc[i] + 8) >> 4;
      rec[(size_t)i * stride + t] = (uint16_t)clampi((int)G.pred[i * N + t] + v, 0, maxv);
    }
  }
  return eob;
}

// grid: (tiles, 2 [luma chain, chroma chain], frames); one warp per CTA
__global__ void __launch_bounds__(32) intra_recon_kernel(const IntraLaunch P) {
  __shared__ ReconGroup G[2];
  const Av1bGeom& g = P.g;
  const int lane = threadIdx.x;
  const int tile = blockIdx.x, chroma = blockIdx.y, frame = blockIdx.z;
  const int tr = tile / g.tile_cols, tc = tile % g.tile_cols;
  TileCtx T;
  T.mi_row_start = g.tile_row_start_sb[tr] * 16; T.mi_row_end = min(g.tile_row_start_sb[tr + 1] * 16, g.mi_rows);
  T.mi_col_start = g.tile_col_start_sb[tc] * 16; T.mi_col_end = min(g.tile_col_start_sb[tc + 1] * 16, g.mi_cols);
  const uint8_t* pmap = P.part_map + (size_t)frame * P.map_elems;
  Av1bBlockInfo* blocks = P.blocks + (size_t)frame * P.map_elems;
  if (lane < 28) { G[0].sw[lane] = tbl::smooth_weights[lane]; G[1].sw[lane] = tbl::smooth_weights[lane]; }
  __syncwarp();
  for (int sb_r = T.mi_row_start; sb_r < T.mi_row_end; sb_r += 16) {
    for (int sb_c = T.mi_col_start; sb_c < T.mi_col_end; sb_c += 16) {
      for (int u = 0; u < 64; u++) {
        const int x8 = (u & 1) | ((u >> 1) & 2) | ((u >> 2) & 4);
        const int y8 = ((u >> 1) & 1) | ((u >> 2) & 2) | ((u >> 3) & 4);
        const int mi_r = sb_r + 2 * y8, mi_c = sb_c + 2 * x8;
        if (mi_r >= g.mi_rows || mi_c >= g.mi_cols) continue;
        const size_t bi = (size_t)(mi_r >> 1) * g.w8 + (mi_c >> 1);
        const int bl = pmap[bi];
        const int n8 = 1 << (bl - 3);
        if ((x8 | y8) & (n8 - 1)) continue;
        const int ha = mi_r > T.mi_row_start, hl = mi_c > T.mi_col_start;
        const int ss = chroma, n = 1 << (bl - ss);
        const int x = (mi_c * 4) >> ss, y = (mi_r * 4) >> ss;
        int har, hbl;
        neighbour_avail(g, T, sb_r, sb_c, ss, (mi_c - sb_c) >> ss, (mi_r - sb_r) >> ss, n >> 2, &har, &hbl);
        int eob = 0;
        if (!chroma) {
          const int mode = blocks[bi].y_mode;
          if (n == 16) { if (lane < 16) eob = intra_tb<16>(P, frame, 0, x, y, mode, T_DCT, T_DCT, ha, hl, har, hbl, lane, 0xFFFFu, G[0]); }
          else { if (lane < 8) eob = intra_tb<8>(P, frame, 0, x, y, mode, T_DCT, T_DCT, ha, hl, har, hbl, lane, 0xFFu, G[0]); }
          eob = __shfl_sync(0xffffffffu, eob, 0);
          for (int o = lane; o < n8 * n8; o += 32) blocks[bi + (size_t)(o / n8) * g.w8 + o % n8].eob[0] = (uint16_t)eob;
        } else {
          const int mode = blocks[bi].uv_mode;
          const int vt = f_mode_vt[mode], ht = f_mode_ht[mode];
          // U on the first group of n lanes, V on the second
          if (n == 8) { if (lane < 16) eob = intra_tb<8>(P, frame, 1 + (lane >> 3), x, y, mode, vt, ht, ha, hl, har, hbl, lane & 7, 0xFFu << (lane & 8), G[lane >> 3]); }
          else { if (lane < 8) eob = intra_tb<4>(P, frame, 1 + (lane >> 2), x, y, mode, vt, ht, ha, hl, har, hbl, lane & 3, 0xFu << (lane & 4), G[lane >> 2]); }
          const int eu = __shfl_sync(0xffffffffu, eob, 0), ev = __shfl_sync(0xffffffffu, eob, n == 8 ? 8 : 4);
          for (int o = lane; o < n8 * n8; o += 32) {
            Av1bBlockInfo& b = blocks[bi + (size_t)(o / n8) * g.w8 + o % n8];
            b.eob[1] = (uint16_t)eu; b.eob[2] = (uint16_t)ev;
          }
        }
        __syncwarp();
      }
    }
  }
}

__global__ void intra_finish_kernel(Av1bBlockInfo* blocks, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Av1bBlockInfo& b = blocks[i];
  if (b.blk_log2) b.skip = (b.eob[0] | b.eob[1] | b.eob[2]) == 0;
}

}  // namespace

// Fast key-frame path: part_map must hold only 16x16 blocks (8x8 at the picture edge).
cudaError_t launch_intra_fast(const IntraLaunch& p, int n_frames, cudaStream_t s) {
  dim3 g1(p.g.sb_cols, p.g.sb_rows, n_frames);
  intra_mode_kernel<<<g1, 256, 0, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  dim3 g2(p.g.tile_cols * p.g.tile_rows, 2, n_frames);
  intra_recon_kernel<<<g2, 32, 0, s>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const size_t n = p.map_elems * n_frames;
  intra_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p.blocks, n);
  return cudaGetLastError();
}

}  // namespace av1b
