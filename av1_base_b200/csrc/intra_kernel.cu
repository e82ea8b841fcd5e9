// Closed-loop intra encode of AV1 key frames on the device: prediction-mode decision (4x4-Hadamard
// SATD with warp shuffles), forward transform, quantisation, normative dequantisation + inverse
// transform + reconstruction.  One CTA owns one tile of one frame and walks its superblocks in raster
// order and their blocks in decode (Z) order, because intra prediction reads reconstructed
// neighbours; tiles of a frame and frames of a batch are independent and fill the grid
// (SURVEY.md 7 "serial dependencies vs GPU width").
//
// Replaces arithmetic the reference delegates to av1an + SVT-AV1
// (/root/reference/crates/daemon/src/encode/av1an.rs:126-139).
// Integer pipes only; the inverse butterflies are the normative ones (av1_inv_txfm1d.h, pinned
// against libaom's av1_idct*/av1_iadst*).  Decisions are defined by oracle/av1_oracle.cpp
// (IntraEnc::block) and must match it bit for bit.
#include <cuda_runtime.h>
#include <stdint.h>
#include "av1_inv_txfm1d.h"
#include "av1_tables_dev.cuh"
#include "kernels.cuh"

namespace av1b {
using namespace av1tx;

namespace {

constexpr int kThreads = 256;
constexpr int kNumCand = 13;
enum { T_DCT = 0, T_ADST = 1, T_FLIP = 2, T_IDT = 3 };

__constant__ uint8_t c_vtype[16] = {T_DCT, T_ADST, T_DCT, T_ADST, T_FLIP, T_DCT, T_FLIP, T_ADST, T_FLIP,
                                    T_IDT, T_DCT, T_IDT, T_ADST, T_IDT, T_FLIP, T_IDT};
__constant__ uint8_t c_htype[16] = {T_DCT, T_DCT, T_ADST, T_ADST, T_DCT, T_FLIP, T_FLIP, T_FLIP, T_ADST,
                                    T_IDT, T_IDT, T_DCT, T_IDT, T_ADST, T_IDT, T_FLIP};
__constant__ uint8_t c_mode_to_txfm[14] = {AV1B_DCT_DCT, AV1B_ADST_DCT, AV1B_DCT_ADST, AV1B_DCT_DCT, AV1B_ADST_ADST,
                                           AV1B_ADST_DCT, AV1B_DCT_ADST, AV1B_DCT_ADST, AV1B_ADST_DCT, AV1B_ADST_ADST,
                                           AV1B_ADST_DCT, AV1B_DCT_ADST, AV1B_ADST_ADST, AV1B_DCT_DCT};
__constant__ uint8_t c_cand[kNumCand] = {AV1B_DC_PRED, AV1B_V_PRED, AV1B_H_PRED, AV1B_PAETH_PRED, AV1B_SMOOTH_PRED,
                                         AV1B_SMOOTH_V_PRED, AV1B_SMOOTH_H_PRED, AV1B_D45_PRED, AV1B_D135_PRED,
                                         AV1B_D113_PRED, AV1B_D157_PRED, AV1B_D203_PRED, AV1B_D67_PRED};

struct Smem {
  uint16_t above[2][136];      // [slot][1 + i], i = -1 .. 2n-1
  uint16_t left[2][136];
  uint16_t src[64 * 64];       // luma block, or U at [0], V at [1024]
  uint16_t pred[64 * 64];
  int16_t resid[64 * 64];
  int32_t bufT[32 * 65];       // forward: column-pass output; inverse: row-pass output (stride n+1)
  int32_t bufC[32 * 32];       // coefficients -> dequantised coefficients
  uint8_t decoded[3][19][19];  // BlockDecoded flags of the current superblock (spec 7.3.? / 5.11.3)
  int cost[kNumCand];
  int dcval[2];
  int eob[3];
  int best;
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ const int16_t* fwd_matrix(int t, int n) {
  if (t == T_DCT) {
    switch (n) {
      case 4: return &tbl::fwd_dct4[0][0];
      case 8: return &tbl::fwd_dct8[0][0];
      case 16: return &tbl::fwd_dct16[0][0];
      case 32: return &tbl::fwd_dct32[0][0];
      default: return &tbl::fwd_dct64[0][0];
    }
  }
  switch (n) {
    case 4: return &tbl::fwd_adst4[0][0];
    case 8: return &tbl::fwd_adst8[0][0];
    default: return &tbl::fwd_adst16[0][0];
  }
}

// ---- edge preparation (spec 7.11.2) ------------------------------------------------------------
__device__ void build_edges(const uint16_t* rec, int stride, int x, int y, int n, int ha, int hl, int har,
                            int hbl, int max_x, int max_y, int bd, uint16_t* above, uint16_t* left) {
  const int base = 1 << (bd - 1);
  const int tid = threadIdx.x;
  // above[1 + i], i = 0 .. 2n-1 ; left likewise ; element 0 is the top-left sample
  for (int i = tid; i < 2 * n; i += kThreads) {
    uint16_t a, l;
    if (ha) {
      const int xx = (i < n || har) ? min(max_x, x + i) : min(max_x, x + n - 1);
      a = rec[(size_t)(y - 1) * stride + xx];
    } else {
      a = hl ? rec[(size_t)y * stride + x - 1] : (uint16_t)(base - 1);
    }
    if (hl) {
      const int yy = (i < n || hbl) ? min(max_y, y + i) : min(max_y, y + n - 1);
      l = rec[(size_t)yy * stride + x - 1];
    } else {
      l = ha ? rec[(size_t)(y - 1) * stride + x] : (uint16_t)(base + 1);
    }
    above[1 + i] = a;
    left[1 + i] = l;
  }
  if (tid == 0) {
    uint16_t tl;
    if (ha && hl) tl = rec[(size_t)(y - 1) * stride + x - 1];
    else if (ha) tl = rec[(size_t)(y - 1) * stride + x];
    else if (hl) tl = rec[(size_t)y * stride + x - 1];
    else tl = (uint16_t)base;
    above[0] = tl;
    left[0] = tl;
  }
}

// one predicted sample; A = above + 1, L = left + 1 (A[-1] = L[-1] = top-left)
__device__ __forceinline__ int pred_px(int mode, int i, int j, int n, const uint16_t* A, const uint16_t* L,
                                       int dcv) {
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
      const uint8_t* w = tbl::smooth_weights + n - 4;
      const int s = w[i] * A[j] + (256 - w[i]) * L[n - 1] + w[j] * L[i] + (256 - w[j]) * A[n - 1];
      return (s + 256) >> 9;
    }
    case AV1B_SMOOTH_V_PRED: {
      const uint8_t* w = tbl::smooth_weights + n - 4;
      return (w[i] * A[j] + (256 - w[i]) * L[n - 1] + 128) >> 8;
    }
    case AV1B_SMOOTH_H_PRED: {
      const uint8_t* w = tbl::smooth_weights + n - 4;
      return (w[j] * L[i] + (256 - w[j]) * A[n - 1] + 128) >> 8;
    }
    default: break;
  }
  // diagonal modes, angle delta 0 (no edge filter / upsampling: enable_intra_edge_filter = 0)
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

// pixel q of an n x n block in 4x4-tile-major order (so that a 16-lane group holds one 4x4 tile)
__device__ __forceinline__ void tile_major(int q, int ln, int& i, int& j) {
  const int t = q >> 4, w = q & 15, tpr = ln - 2;   // tiles per row = n/4 = 1 << (ln-2)
  const int ty = t >> tpr, tx = t & ((1 << tpr) - 1);
  i = ty * 4 + (w >> 2);
  j = tx * 4 + (w & 3);
}

// ---- inverse transform helpers -----------------------------------------------------------------
template <int N>
__device__ __forceinline__ void inv_1d(int t, int32_t* x, int range) {
  if (t == T_DCT) {
    idct<N>(x, range);
  } else if (t == T_IDT) {
    if constexpr (N <= 32) iidentity<N>(x, range);
  } else {
    if constexpr (N == 4) iadst4(x, range);
    else if constexpr (N == 8) iadst8(x, range);
    else if constexpr (N == 16) iadst16(x, range);
  }
}

// Row pass: thread r handles row r of the (CH x CW) dequantised block in bufC, writes N values to bufT
template <int N>
__device__ void inv_rows(const int32_t* bufC, int32_t* bufT, int ht, int row_shift, int row_range) {
  constexpr int C = N < 32 ? N : 32;
  const int r = threadIdx.x;
  if (r < C) {
    int32_t x[N];
#pragma unroll
    for (int j = 0; j < N; j++) x[j] = j < C ? sat(bufC[r * C + j], row_range) : 0;
    inv_1d<N>(ht, x, row_range);
#pragma unroll
    for (int j = 0; j < N; j++) {
      int32_t v = x[ht == T_FLIP ? N - 1 - j : j];
      if (row_shift > 0) v = (v + (1 << (row_shift - 1))) >> row_shift;
      bufT[r * (N + 1) + j] = v;
    }
  }
}

// Column pass: thread c handles column c, adds to the prediction and writes the reconstruction
template <int N>
__device__ void inv_cols(const int32_t* bufT, const uint16_t* pred, uint16_t* rec, int rstride, int vt,
                         int col_range, int maxv) {
  constexpr int C = N < 32 ? N : 32;
  const int c = threadIdx.x;
  if (c < N) {
    int32_t x[N];
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = i < C ? sat(bufT[i * (N + 1) + c], col_range) : 0;
    inv_1d<N>(vt, x, col_range);
#pragma unroll
    for (int i = 0; i < N; i++) {
      const int32_t v = (x[vt == T_FLIP ? N - 1 - i : i] + 8) >> 4;
      rec[(size_t)i * rstride + c] = (uint16_t)clampi((int)pred[i * N + c] + v, 0, maxv);
    }
  }
}

// ---- one plane of one block: residual -> forward -> quant -> (dequant -> inverse -> recon) -----
__device__ void code_plane(Smem& sm, const IntraLaunch& P, int plane, int slot, int n, int ln, int mode,
                           int tx_type, int x, int y, uint16_t* rec, int16_t* coef, int stride, int dcv) {
  const int tid = threadIdx.x;
  const uint16_t* A = sm.above[slot] + 1;
  const uint16_t* L = sm.left[slot] + 1;
  const uint16_t* src = sm.src + slot * 1024;
  const int npx = n * n;
  // prediction + residual
  for (int q = tid; q < npx; q += kThreads) {
    const int i = q >> ln, j = q & (n - 1);
    const int p = pred_px(mode, i, j, n, A, L, dcv);
    sm.pred[q] = (uint16_t)p;
    sm.resid[q] = (int16_t)((int)src[q] - p);
  }
  if (tid == 0) sm.eob[plane] = 0;
  __syncthreads();
  const int vt = c_vtype[tx_type], ht = c_htype[tx_type];
  const int cn = n < 32 ? n : 32, lcn = ln < 5 ? ln : 5;
  // forward, column pass: T[k][j] = (sum_i Fv[k][i] * 4*r[i][j] + 2048) >> 12
  {
    const int16_t* F = fwd_matrix(vt, n);
    for (int o = tid; o < cn * n; o += kThreads) {
      const int k = o >> ln, j = o & (n - 1);
      const int jj = ht == T_FLIP ? n - 1 - j : j;
      int32_t acc = 0;
      if (vt == T_FLIP) {
        for (int i = 0; i < n; i++) acc += F[k * n + i] * ((int32_t)sm.resid[(n - 1 - i) * n + jj] * 4);
      } else {
        for (int i = 0; i < n; i++) acc += F[k * n + i] * ((int32_t)sm.resid[i * n + jj] * 4);
      }
      sm.bufT[k * (n + 1) + j] = (acc + 2048) >> 12;
    }
  }
  __syncthreads();
  // forward, row pass + scaling + quantisation (+ normative dequantisation)
  {
    const int16_t* F = fwd_matrix(ht, n);
    const int row_shift_tab = ln == 2 ? 0 : ln == 3 ? 1 : 2;     // Transform_Row_Shift for squares
    const int sh = 24 + 2 * ln - row_shift_tab - 4;
    const int s = (npx > 256) + (npx > 1024);
    const int lim = (1 << (7 + P.bit_depth)) - 1;
    const int16_t* iscan = cn == 4 ? tbl::iscan_default_4 : cn == 8 ? tbl::iscan_default_8
                           : cn == 16 ? tbl::iscan_default_16 : tbl::iscan_default_32;
    // the cn x cn matrix of this plane's level (av1_qm_tables.h: 4x4 at 0, 8x8 at 16, 16x16 at 80, 32x32 at 336) or flat
    const uint8_t* qmw = P.qm[plane > 0] ? P.qm[plane > 0] + (cn == 4 ? 0 : cn == 8 ? 16 : cn == 16 ? 80 : 336) : nullptr;
    for (int o = tid; o < cn * cn; o += kThreads) {
      // lanes run over k (rows) so that F[l][j] is warp-uniform
      const int k = o & (cn - 1), l = o >> lcn;
      int64_t acc = 0;
      for (int j = 0; j < n; j++) acc += (int64_t)F[l * n + j] * sm.bufT[k * (n + 1) + j];
      const int32_t c = (int32_t)((acc * 4096 + ((int64_t)1 << (sh - 1))) >> sh);
      int dqv = (k | l) ? P.ac_q : P.dc_q;
      if (qmw) dqv = (dqv * (int)qmw[k * cn + l] + 16) >> 5;   // quantisation matrix: Round2(q * weight, 5), spec 7.12.3
      const uint32_t a = (uint32_t)(c < 0 ? -c : c) << s;
      uint32_t lv = (a + (uint32_t)((dqv * P.quant_rnd) >> 7)) / (uint32_t)dqv;
      if (lv > 32767u) lv = 32767u;
      int32_t d = (int32_t)(((lv * (uint32_t)dqv) & 0xFFFFFFu) >> s);
      if (d > lim) d = lim;
      if (c < 0) d = -d;
      sm.bufC[k * cn + l] = d;
      coef[k * cn + l] = (int16_t)(c < 0 ? -(int32_t)lv : (int32_t)lv);
      if (lv) atomicMax(&sm.eob[plane], (int)iscan[k * cn + l] + 1);
    }
    // a 64x64 block owns 4096 elements of the level plane and codes 32x32 of them: the rest reads as zero, whatever the
    // batch slot held before (the symbol streams of a frame are compared as a whole in debug mode)
    for (int o = cn * cn + tid; o < npx; o += kThreads) coef[o] = 0;
  }
  __syncthreads();
  const int eob = sm.eob[plane];
  const int maxv = (1 << P.bit_depth) - 1;
  if (eob == 0) {
    for (int q = tid; q < npx; q += kThreads) rec[(size_t)(q >> ln) * stride + (q & (n - 1))] = sm.pred[q];
  } else {
    const int row_range = P.bit_depth + 8, col_range = max(P.bit_depth + 6, 16);
    const int row_shift = ln == 2 ? 0 : ln == 3 ? 1 : 2;
    switch (n) {
      case 4: inv_rows<4>(sm.bufC, sm.bufT, ht, row_shift, row_range); break;
      case 8: inv_rows<8>(sm.bufC, sm.bufT, ht, row_shift, row_range); break;
      case 16: inv_rows<16>(sm.bufC, sm.bufT, ht, row_shift, row_range); break;
      case 32: inv_rows<32>(sm.bufC, sm.bufT, ht, row_shift, row_range); break;
      default: inv_rows<64>(sm.bufC, sm.bufT, ht, row_shift, row_range); break;
    }
    __syncthreads();
    switch (n) {
      case 4: inv_cols<4>(sm.bufT, sm.pred, rec, stride, vt, col_range, maxv); break;
      case 8: inv_cols<8>(sm.bufT, sm.pred, rec, stride, vt, col_range, maxv); break;
      case 16: inv_cols<16>(sm.bufT, sm.pred, rec, stride, vt, col_range, maxv); break;
      case 32: inv_cols<32>(sm.bufT, sm.pred, rec, stride, vt, col_range, maxv); break;
      default: inv_cols<64>(sm.bufT, sm.pred, rec, stride, vt, col_range, maxv); break;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads) intra_encode_kernel(const IntraLaunch P) {
  __shared__ Smem sm;
  const Av1bGeom& g = P.g;
  const int tid = threadIdx.x, lane = tid & 31;
  const int tile = blockIdx.x, frame = blockIdx.y;
  const int tr = tile / g.tile_cols, tc = tile % g.tile_cols;
  const int mi_row_start = g.tile_row_start_sb[tr] * 16, mi_row_end = min(g.tile_row_start_sb[tr + 1] * 16, g.mi_rows);
  const int mi_col_start = g.tile_col_start_sb[tc] * 16, mi_col_end = min(g.tile_col_start_sb[tc + 1] * 16, g.mi_cols);
  const uint16_t* srcp[3];
  uint16_t* recp[3];
  int16_t* coefp[3];
  for (int p = 0; p < 3; p++) {
    srcp[p] = P.src[p] + (size_t)frame * P.plane_elems[p];
    recp[p] = P.rec[p] + (size_t)frame * P.plane_elems[p];
    coefp[p] = P.coef[p] + (size_t)frame * P.plane_elems[p];
  }
  Av1bBlockInfo* blocks = P.blocks + (size_t)frame * P.map_elems;
  const uint8_t* pmap = P.part_map + (size_t)frame * P.map_elems;
  const int bd = P.bit_depth;

  for (int sb_r = mi_row_start; sb_r < mi_row_end; sb_r += 16) {
    for (int sb_c = mi_col_start; sb_c < mi_col_end; sb_c += 16) {
      // clear_block_decoded_flags (spec 5.11.3)
      for (int o = tid; o < 3 * 19 * 19; o += kThreads) {
        const int p = o / 361, rem = o - p * 361, yy = rem / 19 - 1, xx = rem % 19 - 1;
        const int ss = p > 0, nn = 16 >> ss;
        const int sbw4 = (mi_col_end - sb_c) >> ss, sbh4 = (mi_row_end - sb_r) >> ss;
        uint8_t v = 0;
        if (yy <= nn && xx <= nn) {
          if (yy < 0 && xx < sbw4) v = 1;
          else if (xx < 0 && yy < sbh4) v = 1;
          if (yy == nn && xx == -1) v = 0;
        }
        sm.decoded[p][yy + 1][xx + 1] = v;
      }
      __syncthreads();
      for (int u = 0; u < 64; u++) {
        const int x8 = (u & 1) | ((u >> 1) & 2) | ((u >> 2) & 4);
        const int y8 = ((u >> 1) & 1) | ((u >> 2) & 2) | ((u >> 3) & 4);
        const int mi_r = sb_r + 2 * y8, mi_c = sb_c + 2 * x8;
        if (mi_r >= g.mi_rows || mi_c >= g.mi_cols) continue;
        const int bl = pmap[(mi_r >> 1) * g.w8 + (mi_c >> 1)];
        const int n8 = 1 << (bl - 3);
        if ((x8 | y8) & (n8 - 1)) continue;
        // ---------------- one block ----------------
        int y_mode = 0, uv_mode = 0;
        for (int pass = 0; pass < 2; pass++) {
          const int ss = pass;
          const int n = min(1 << (bl - ss), pass ? 32 : 64);
          const int ln = 31 - __clz(n);
          const int nplanes = pass ? 2 : 1;
          const int x = (mi_c * 4) >> ss, y = (mi_r * 4) >> ss;
          const int ha = mi_r > mi_row_start, hl = mi_c > mi_col_start;
          const int x4 = (mi_c - sb_c) >> ss, y4 = (mi_r - sb_r) >> ss, n4 = n >> 2;
          const int max_x = ((g.mi_cols * 4) >> ss) - 1, max_y = ((g.mi_rows * 4) >> ss) - 1;
          // open-loop decision: the candidate predictors are built from the SOURCE picture's neighbours
          for (int k = 0; k < nplanes; k++) {
            const int p = pass + k;
            const int har = sm.decoded[p][y4][x4 + n4 + 1], hbl = sm.decoded[p][y4 + n4 + 1][x4];
            build_edges(srcp[p], g.stride[p], x, y, n, ha, hl, har, hbl, max_x, max_y, bd, sm.above[k], sm.left[k]);
            const uint16_t* s = srcp[p] + (size_t)y * g.stride[p] + x;
            for (int q = tid; q < n * n; q += kThreads) sm.src[k * 1024 + q] = s[(size_t)(q >> ln) * g.stride[p] + (q & (n - 1))];
          }
          if (tid < kNumCand) sm.cost[tid] = 0;
          __syncthreads();
          // DC values: warp k sums plane k's edges
          if ((tid >> 5) < nplanes) {
            const int k = tid >> 5;
            const uint16_t* A = sm.above[k] + 1;
            const uint16_t* L = sm.left[k] + 1;
            int s = 0;
            for (int i = lane; i < n; i += 32) s += (ha ? A[i] : 0) + (hl ? L[i] : 0);
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            int v;
            if (ha && hl) v = (s + n) >> (ln + 1);
            else if (ha || hl) v = (s + (n >> 1)) >> ln;
            else v = 1 << (bd - 1);
            if (lane == 0) sm.dcval[k] = v;
          }
          __syncthreads();
          // mode decision: SATD (sum |4x4 Hadamard|) of source - prediction, every candidate
          {
            const int npx = n * n;
            for (int m = 0; m < kNumCand; m++) {
              const int mode = c_cand[m];
              int acc = 0;
              for (int k = 0; k < nplanes; k++) {
                const uint16_t* A = sm.above[k] + 1;
                const uint16_t* L = sm.left[k] + 1;
                const int dcv = sm.dcval[k];
                for (int base = tid - lane; base < npx; base += kThreads) {
                  const int q = base + lane;
                  int v = 0;
                  if (q < npx) {
                    int i, j;
                    tile_major(q, ln, i, j);
                    v = (int)sm.src[k * 1024 + i * n + j] - pred_px(mode, i, j, n, A, L, dcv);
                  }
                  // 4x4 Hadamard across the 16-lane group: lanes = (row << 2 | col)
#pragma unroll
                  for (int msk = 1; msk <= 8; msk <<= 1) {
                    const int o = __shfl_xor_sync(0xffffffffu, v, msk);
                    v = (lane & msk) ? o - v : v + o;
                  }
                  acc += abs(v);
                }
              }
              for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
              if (lane == 0 && acc) atomicAdd(&sm.cost[m], acc);
            }
          }
          __syncthreads();
          if (tid == 0) {
            int best = 0, bc = sm.cost[0];
            for (int m = 1; m < kNumCand; m++) if (sm.cost[m] < bc) { bc = sm.cost[m]; best = m; }
            sm.best = c_cand[best];
          }
          __syncthreads();
          const int mode = sm.best;
          if (pass == 0) y_mode = mode; else uv_mode = mode;
          // closed loop: the coded prediction uses the reconstructed neighbours
          for (int k = 0; k < nplanes; k++) {
            const int p = pass + k;
            const int har = sm.decoded[p][y4][x4 + n4 + 1], hbl = sm.decoded[p][y4 + n4 + 1][x4];
            build_edges(recp[p], g.stride[p], x, y, n, ha, hl, har, hbl, max_x, max_y, bd, sm.above[k], sm.left[k]);
          }
          __syncthreads();
          if ((tid >> 5) < nplanes) {
            const int k = tid >> 5;
            const uint16_t* A = sm.above[k] + 1;
            const uint16_t* L = sm.left[k] + 1;
            int s = 0;
            for (int i = lane; i < n; i += 32) s += (ha ? A[i] : 0) + (hl ? L[i] : 0);
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            int v;
            if (ha && hl) v = (s + n) >> (ln + 1);
            else if (ha || hl) v = (s + (n >> 1)) >> ln;
            else v = 1 << (bd - 1);
            if (lane == 0) sm.dcval[k] = v;
          }
          __syncthreads();
          for (int k = 0; k < nplanes; k++) {
            const int p = pass + k;
            int tx_type = AV1B_DCT_DCT;
            if (p > 0 && n < 32) tx_type = c_mode_to_txfm[mode];
            code_plane(sm, P, p, k, n, ln, mode, tx_type, x, y, recp[p] + (size_t)y * g.stride[p] + x,
                       coefp[p] + av1b_coef_offset(g.sb_cols, p, x, y), g.stride[p], sm.dcval[k]);
          }
        }
        // publish the block's side information and mark it decoded
        {
          Av1bBlockInfo info;
          info.blk_log2 = (uint8_t)bl; info.y_mode = (uint8_t)y_mode; info.uv_mode = (uint8_t)uv_mode;
          info.skip = (sm.eob[0] | sm.eob[1] | sm.eob[2]) == 0;
          info.angle_y = 0; info.angle_uv = 0; info.tx_type_y = AV1B_DCT_DCT; info.cfl_alpha_u = 0;
          info.eob[0] = (uint16_t)sm.eob[0]; info.eob[1] = (uint16_t)sm.eob[1]; info.eob[2] = (uint16_t)sm.eob[2];
          info.cfl_alpha_v = 0; info.is_inter = 0; info.mv[0] = 0; info.mv[1] = 0;
          for (int o = tid; o < n8 * n8; o += kThreads) {
            const int yy = o / n8, xx = o % n8;
            blocks[((mi_r >> 1) + yy) * g.w8 + (mi_c >> 1) + xx] = info;
          }
          const int n4l = 1 << (bl - 2);
          for (int o = tid; o < n4l * n4l; o += kThreads) {
            const int yy = o / n4l, xx = o % n4l;
            sm.decoded[0][(mi_r - sb_r) + yy + 1][(mi_c - sb_c) + xx + 1] = 1;
            if (!((yy | xx) & 1)) {
              sm.decoded[1][((mi_r - sb_r) >> 1) + (yy >> 1) + 1][((mi_c - sb_c) >> 1) + (xx >> 1) + 1] = 1;
              sm.decoded[2][((mi_r - sb_r) >> 1) + (yy >> 1) + 1][((mi_c - sb_c) >> 1) + (xx >> 1) + 1] = 1;
            }
          }
        }
        __syncthreads();
      }
    }
  }
}

__global__ void partition_fixed_kernel(Av1bGeom g, int blk_log2, uint8_t* map, int n_frames) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = g.w8 * g.h8;
  if (idx >= per * n_frames) return;
  const int u = idx % per, y = u / g.w8, x = u % g.w8;
  int bl = blk_log2;
  while (bl > 3) {
    const int n8 = 1 << (bl - 3), x0 = x & ~(n8 - 1), y0 = y & ~(n8 - 1);
    if (x0 + n8 <= g.w8 && y0 + n8 <= g.h8) break;
    bl--;
  }
  map[idx] = (uint8_t)bl;
}

// Key-frame partition by smoothness (oracle: orc_partition_smooth).  One CTA per 64x64 superblock, thread (i, j) owns the
// 4x4 luma box (i, j): 16x16 box sums, the plane through the block mean with slopes from the half sums, and the largest
// deviation of a box from it decide 64x64 / 32x32 / fixed 16x16.
__global__ void __launch_bounds__(256) partition_smooth_kernel(Av1bGeom g, const uint16_t* __restrict__ src_y, size_t plane_elems,
                                                               size_t map_elems, int thr, uint8_t* map) {
  __shared__ int s_sum[5][5];    // [0]: the 64x64 block, [1 + q]: 32x32 quadrant q; S, SL, SR, ST, SB
  __shared__ int s_bad[5];
  const int tid = threadIdx.x, i = tid >> 4, j = tid & 15;
  const int frame = blockIdx.z, x0 = blockIdx.x * 64, y0 = blockIdx.y * 64;
  const uint16_t* src = src_y + (size_t)frame * plane_elems;
  if (tid < 25) s_sum[tid / 5][tid % 5] = 0;
  if (tid < 5) s_bad[tid] = 0;
  __syncthreads();
  int box = 0;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uint2 v = *reinterpret_cast<const uint2*>(src + (size_t)(y0 + 4 * i + r) * g.stride[0] + x0 + 4 * j);
    box += (int)((v.x & 0xFFFF) + (v.x >> 16) + (v.y & 0xFFFF) + (v.y >> 16));
  }
  const int q = (i >> 3) * 2 + (j >> 3), i8 = i & 7, j8 = j & 7;
  atomicAdd(&s_sum[0][0], box); atomicAdd(&s_sum[0][j < 8 ? 1 : 2], box); atomicAdd(&s_sum[0][i < 8 ? 3 : 4], box);
  atomicAdd(&s_sum[1 + q][0], box); atomicAdd(&s_sum[1 + q][j8 < 4 ? 1 : 2], box); atomicAdd(&s_sum[1 + q][i8 < 4 ? 3 : 4], box);
  __syncthreads();
  {
    const int* t = s_sum[0];
    const int d = 4096 * box - (16 * t[0] + (2 * j - 15) * 2 * (t[2] - t[1]) + (2 * i - 15) * 2 * (t[4] - t[3]));
    if (abs(d) > thr * 4096) s_bad[0] = 1;
    const int* u = s_sum[1 + q];
    const int e = 512 * box - (8 * u[0] + (2 * j8 - 7) * 2 * (u[2] - u[1]) + (2 * i8 - 7) * 2 * (u[4] - u[3]));
    if (abs(e) > thr * 512) s_bad[1 + q] = 1;
  }
  __syncthreads();
  if (tid < 64) {
    const int ux = blockIdx.x * 8 + (tid & 7), uy = blockIdx.y * 8 + (tid >> 3);
    if (ux < g.w8 && uy < g.h8) {
      int bl = 4;
      while (bl > 3) {
        const int n8 = 1 << (bl - 3), xa = ux & ~(n8 - 1), ya = uy & ~(n8 - 1);
        if (xa + n8 <= g.w8 && ya + n8 <= g.h8) break;
        bl--;
      }
      const int qq = ((tid >> 3) >> 2) * 2 + ((tid & 7) >> 2);
      const int qx = x0 + (qq & 1) * 32, qy = y0 + (qq >> 1) * 32;
      if (x0 + 64 <= g.width && y0 + 64 <= g.height && !s_bad[0]) bl = 6;
      else if (qx + 32 <= g.width && qy + 32 <= g.height && !s_bad[1 + qq]) bl = 5;
      map[(size_t)frame * map_elems + (size_t)uy * g.w8 + ux] = (uint8_t)bl;
    }
  }
}

}  // namespace

cudaError_t launch_partition_smooth(const Av1bGeom& g, const uint16_t* src_y, size_t plane_elems, size_t map_elems, int thr,
                                    uint8_t* map, int n_frames, cudaStream_t s) {
  dim3 grid(g.sb_cols, g.sb_rows, n_frames);
  // a box sum deviates from the plane by less than 2^15 whatever the picture: larger thresholds mean "always smooth"
  // (and thr * 4096 stays inside 32 bits)
  thr = thr > (1 << 17) ? (1 << 17) : thr;
  partition_smooth_kernel<<<grid, 256, 0, s>>>(g, src_y, plane_elems, map_elems, thr, map);
  return cudaGetLastError();
}

cudaError_t launch_partition_fixed(const Av1bGeom& g, int blk_log2, uint8_t* map, int n_frames, cudaStream_t s) {
  const int total = g.w8 * g.h8 * n_frames;
  partition_fixed_kernel<<<(total + 255) / 256, 256, 0, s>>>(g, blk_log2, map, n_frames);
  return cudaGetLastError();
}

cudaError_t launch_intra_encode(const IntraLaunch& p, int n_frames, cudaStream_t s) {
  dim3 grid(p.g.tile_cols * p.g.tile_rows, n_frames);
  intra_encode_kernel<<<grid, kThreads, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace av1b
