// Inter frame encode on the device: motion-compensated prediction (normative 8-tap / 4-tap
// interpolation, spec 7.11.3.4), residual, forward DCT, quantisation, normative dequantisation +
// inverse DCT + reconstruction.  Blocks of an inter frame do not depend on each other, so the whole
// frame runs in parallel: one CTA per 64x64 superblock, and inside it every transform block is owned
// by a group of N threads (N = transform size: 16 luma / 8 chroma for 16x16 blocks, 8 / 4 for the
// 8x8 blocks at the picture edge).  A group never leaves its warp, so the only synchronisation is
// __syncwarp on the group's lanes.  Thread t of a group owns row t (residual, forward row pass,
// quantiser, inverse row pass) and column t (forward column pass, inverse column pass + store), which
// makes every transform-matrix operand warp-uniform (constant bank) and every shared-memory access
// conflict-free.
//
// Replaces arithmetic the reference delegates to av1an + SVT-AV1
// (/root/reference/crates/daemon/src/encode/av1an.rs:126-139; SURVEY.md 8a rows E4, E5 for inter frames).
// Must match oracle/av1_oracle.cpp orc_encode_inter_frame bit for bit.
#include <cuda_runtime.h>
#include <stdint.h>
#include "av1_inv_txfm1d.h"
#include "av1_tables_dev.cuh"
#include "kernels.cuh"

namespace av1b {
using namespace av1tx;
namespace {

constexpr int kThreads = 256;
constexpr int kPoolBytes = 16 * ((16 + 7) * 17 * 4 + 16 * 16 * 2 + 18 * 18 + 4);   // 16 groups of N = 16 (largest carving)

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

template <int N> struct TxTab;
template <> struct TxTab<4> {
  // forward DCT matrix (av1_fwd_matrices.h) as a function-local constexpr table: after unrolling every entry is an
  // immediate operand instead of a constant-bank load
  static __device__ __forceinline__ int f(int k, int i) {
    constexpr int16_t t[4][4] = {{2896, 2896, 2896, 2896},
                                  {3784, 1567, -1567, -3784},
                                  {2896, -2896, -2896, 2896},
                                  {1567, -3784, 3784, -1567}};
    return t[k][i];
  }
  static __device__ __forceinline__ int iscan(int p) { return tbl::iscan_default_4[p]; }
  static __device__ __forceinline__ int nz_off(int p) { return tbl::nz_map_ctx_offset_4[p]; }
  static constexpr int kLog2 = 2, kRowShift = 0;
};
template <> struct TxTab<8> {
  // forward DCT matrix (av1_fwd_matrices.h) as a function-local constexpr table: after unrolling every entry is an
  // immediate operand instead of a constant-bank load
  static __device__ __forceinline__ int f(int k, int i) {
    constexpr int16_t t[8][8] = {{2896, 2896, 2896, 2896, 2896, 2896, 2896, 2896},
                                  {4017, 3406, 2276, 799, -799, -2276, -3406, -4017},
                                  {3784, 1567, -1567, -3784, -3784, -1567, 1567, 3784},
                                  {3406, -799, -4017, -2276, 2276, 4017, 799, -3406},
                                  {2896, -2896, -2896, 2896, 2896, -2896, -2896, 2896},
                                  {2276, -4017, 799, 3406, -3406, -799, 4017, -2276},
                                  {1567, -3784, 3784, -1567, -1567, 3784, -3784, 1567},
                                  {799, -2276, 3406, -4017, 4017, -3406, 2276, -799}};
    return t[k][i];
  }
  static __device__ __forceinline__ int iscan(int p) { return tbl::iscan_default_8[p]; }
  static __device__ __forceinline__ int nz_off(int p) { return tbl::nz_map_ctx_offset_8[p]; }
  static constexpr int kLog2 = 3, kRowShift = 1;
};
template <> struct TxTab<16> {
  // forward DCT matrix (av1_fwd_matrices.h) as a function-local constexpr table: after unrolling every entry is an
  // immediate operand instead of a constant-bank load
  static __device__ __forceinline__ int f(int k, int i) {
    constexpr int16_t t[16][16] = {{2896, 2896, 2896, 2896, 2896, 2896, 2896, 2896, 2896, 2896, 2896, 2896, 2896, 2896, 2896, 2896},
                                  {4076, 3920, 3612, 3166, 2598, 1931, 1189, 401, -401, -1189, -1931, -2598, -3166, -3612, -3920, -4076},
                                  {4017, 3406, 2276, 799, -799, -2276, -3406, -4017, -4017, -3406, -2276, -799, 799, 2276, 3406, 4017},
                                  {3920, 2598, 401, -1931, -3612, -4076, -3166, -1189, 1189, 3166, 4076, 3612, 1931, -401, -2598, -3920},
                                  {3784, 1567, -1567, -3784, -3784, -1567, 1567, 3784, 3784, 1567, -1567, -3784, -3784, -1567, 1567, 3784},
                                  {3612, 401, -3166, -3920, -1189, 2598, 4076, 1931, -1931, -4076, -2598, 1189, 3920, 3166, -401, -3612},
                                  {3406, -799, -4017, -2276, 2276, 4017, 799, -3406, -3406, 799, 4017, 2276, -2276, -4017, -799, 3406},
                                  {3166, -1931, -3920, 401, 4076, 1189, -3612, -2598, 2598, 3612, -1189, -4076, -401, 3920, 1931, -3166},
                                  {2896, -2896, -2896, 2896, 2896, -2896, -2896, 2896, 2896, -2896, -2896, 2896, 2896, -2896, -2896, 2896},
                                  {2598, -3612, -1189, 4076, -401, -3920, 1931, 3166, -3166, -1931, 3920, 401, -4076, 1189, 3612, -2598},
                                  {2276, -4017, 799, 3406, -3406, -799, 4017, -2276, -2276, 4017, -799, -3406, 3406, 799, -4017, 2276},
                                  {1931, -4076, 2598, 1189, -3920, 3166, 401, -3612, 3612, -401, -3166, 3920, -1189, -2598, 4076, -1931},
                                  {1567, -3784, 3784, -1567, -1567, 3784, -3784, 1567, 1567, -3784, 3784, -1567, -1567, 3784, -3784, 1567},
                                  {1189, -3166, 4076, -3612, 1931, 401, -2598, 3920, -3920, 2598, -401, -1931, 3612, -4076, 3166, -1189},
                                  {799, -2276, 3406, -4017, 4017, -3406, 2276, -799, -799, 2276, -3406, 4017, -4017, 3406, -2276, 799},
                                  {401, -1189, 1931, -2598, 3166, -3612, 3920, -4076, 4076, -3920, 3612, -3166, 2598, -1931, 1189, -401}};
    return t[k][i];
  }
  static __device__ __forceinline__ int iscan(int p) { return tbl::iscan_default_16[p]; }
  static __device__ __forceinline__ int nz_off(int p) { return tbl::nz_map_ctx_offset_16[p]; }
  static constexpr int kLog2 = 4, kRowShift = 2;
};

// One transform block == one prediction block of plane `p` at (x, y), size N x N, owned by the N lanes
// `gmask` of one warp; t = lane index inside the group.  buf: (N+7)*(N+1) int32, pred: N*N uint16.
// QM: quantisation matrices in force (P.qm: the step of every position is weighted, spec 7.12.3); the <false> instance is
// the code as it was without them.
template <int N, bool QM>
__device__ __forceinline__ int code_tb(const InterLaunch& P, int frame, int p, int x, int y, int mv_row, int mv_col, int t,
                                       unsigned gmask, int32_t* buf, uint16_t* pred, uint8_t* lv8, bool active) {
  constexpr int S = N + 1;
  const int ss = p > 0, bd = P.bit_depth;
  const int stride = P.g.stride[p];
  const size_t fo = (size_t)frame * P.plane_elems[p];   // this frame's planes inside the launch's batch buffers
  const int pw = P.g.width >> ss, ph = P.g.height >> ss;
  const uint16_t* ref = P.ref[p];
  // ---------------- prediction ----------------
  const int x16 = (x << 4) + ((2 * mv_col) >> ss), y16 = (y << 4) + ((2 * mv_row) >> ss);
  const int ix = x16 >> 4, iy = y16 >> 4, fx = x16 & 15, fy = y16 & 15;
  // rows that lie inside the picture with their whole window are read as aligned 32-bit words without clamping (an odd
  // start is a funnel shift); N = 4 keeps the sample loads
  const bool interior = N >= 8 && ix >= 4 && iy >= 3 && ix + N + 5 <= pw && iy + N + 4 <= ph;
  if (fx == 0 && fy == 0) {
    if (interior) {
      const int xo = ix & 1;
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(ref + (size_t)(iy + t) * stride + (ix - xo));
      uint32_t w[N / 2 + 1];
#pragma unroll
      for (int k = 0; k < N / 2 + 1; k++) w[k] = rw[k];
      if (xo) {
#pragma unroll
        for (int k = 0; k < N / 2; k++) w[k] = __funnelshift_r(w[k], w[k + 1], 16);
      }
      uint32_t* pw32 = reinterpret_cast<uint32_t*>(pred + t * N);
#pragma unroll
      for (int k = 0; k < N / 2; k++) pw32[k] = w[k];
    } else {
      const uint16_t* rr = ref + (size_t)clampi(iy + t, 0, ph - 1) * stride;
#pragma unroll
      for (int c = 0; c < N; c++) pred[t * N + c] = rr[clampi(ix + c, 0, pw - 1)];
    }
  } else {
    int kx[8], ky[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      kx[k] = N <= 4 ? tbl::sub_pel_filters_4[fx][k] : tbl::sub_pel_filters_8[fx][k];
      ky[k] = N <= 4 ? tbl::sub_pel_filters_4[fy][k] : tbl::sub_pel_filters_8[fy][k];
    }
    for (int r = t; r < N + 7; r += N) {
      int win[N + 7];
      if (interior) {
        const int wx = ix - 3, xo = wx & 1;
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(ref + (size_t)(iy + r - 3) * stride + (wx - xo));
        constexpr int WW = (N + 7 + 2) / 2;   // words that cover N + 7 samples from an even or odd start
        uint32_t w[WW];
#pragma unroll
        for (int k = 0; k < WW; k++) w[k] = rw[k];
        if (xo) {
#pragma unroll
          for (int k = 0; k < WW - 1; k++) w[k] = __funnelshift_r(w[k], w[k + 1], 16);
        }
#pragma unroll
        for (int c = 0; c < N + 7; c++) win[c] = (c & 1) ? (int)(w[c >> 1] >> 16) : (int)(w[c >> 1] & 0xFFFFu);
      } else {
        const uint16_t* rr = ref + (size_t)clampi(iy + r - 3, 0, ph - 1) * stride;
#pragma unroll
        for (int c = 0; c < N + 7; c++) win[c] = rr[clampi(ix + c - 3, 0, pw - 1)];
      }
#pragma unroll
      for (int c = 0; c < N; c++) {
        int s = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) s += kx[k] * win[c + k];
        buf[r * S + c] = (s + 4) >> 3;
      }
    }
    __syncwarp(gmask);
    const int maxv = (1 << bd) - 1;
#pragma unroll
    for (int c = 0; c < N; c++) {
      int s = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) s += ky[k] * buf[(t + k) * S + c];
      pred[t * N + c] = (uint16_t)clampi((s + 1024) >> 11, 0, maxv);
    }
    __syncwarp(gmask);
  }
  // ---------------- residual (row t) -> shared, as the column pass reads columns ----------------
  {
    const uint16_t* sp = P.src[p] + fo + (size_t)(y + t) * stride + x;
#pragma unroll
    for (int c = 0; c < N; c++) buf[t * S + c] = ((int)sp[c] - (int)pred[t * N + c]) * 4;
  }
  // ---------------- early skip (oracle: orc_encode_inter_frame): residual sums to less than dc_q * N / 16 ----------
  {
    int sad4 = 0;
#pragma unroll
    for (int c = 0; c < N; c++) sad4 += abs(buf[t * S + c]);
#pragma unroll
    for (int o = N / 2; o; o >>= 1) sad4 += __shfl_xor_sync(gmask, sad4, o);
    if (sad4 < 4 * ((P.dc_q * N) >> 4)) {
      if (active) {
        int16_t* cz = P.coef[p] + fo + av1b_coef_offset(P.g.sb_cols, p, x, y) + t * N;
#pragma unroll
        for (int l = 0; l < N; l += 4) *reinterpret_cast<uint2*>(cz + l) = make_uint2(0u, 0u);
        uint16_t* rec = P.rec[p] + fo + (size_t)y * stride + x;
#pragma unroll
        for (int i = 0; i < N; i++) rec[(size_t)i * stride + t] = pred[i * N + t];
      }
      __syncwarp(gmask);
      return 0;
    }
  }
  __syncwarp(gmask);
  // ---------------- forward DCT: column pass (thread t = column t), then row pass (thread t = row t) -----
  // F[k][N-1-i] = F[k][i] for even k and -F[k][i] for odd k (checked when the tables were generated): the sums below
  // are the full matrix products term for term, with half the multiplications
  int32_t col[N];
  {
    int32_t se[N / 2], so[N / 2];
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
      const int32_t a = buf[i * S + t], b = buf[(N - 1 - i) * S + t];
      se[i] = a + b; so[i] = a - b;
    }
#pragma unroll
    for (int k = 0; k < N; k++) {
      int32_t acc = 0;
#pragma unroll
      for (int i = 0; i < N / 2; i++) acc += TxTab<N>::f(k, i) * ((k & 1) ? so[i] : se[i]);
      col[k] = (acc + 2048) >> 12;
    }
  }
  __syncwarp(gmask);
#pragma unroll
  for (int k = 0; k < N; k++) buf[k * S + t] = col[k];
  __syncwarp(gmask);
  int eob = 0, sum_abs = 0;
  int16_t* cbase = P.coef[p] + fo + av1b_coef_offset(P.g.sb_cols, p, x, y);
  int16_t* cdst = cbase + t * N;
  constexpr int S8 = N + 2;           // pitch of the capped-magnitude map (two zero guard rows / columns)
  unsigned sign_bits = 0, max_lv = 0;
  {
    int32_t rse[N / 2], rso[N / 2];
#pragma unroll
    for (int j = 0; j < N / 2; j++) {
      const int32_t a = buf[t * S + j], b = buf[t * S + N - 1 - j];
      rse[j] = a + b; rso[j] = a - b;
    }
    constexpr int sh = 24 + 2 * TxTab<N>::kLog2 - TxTab<N>::kRowShift - 4;
    const int lim = (1 << (7 + bd)) - 1;
    const uint32_t rnd_dc = (uint32_t)((P.dc_q * P.quant_rnd) >> 7), rnd_ac = (uint32_t)((P.ac_q * P.quant_rnd) >> 7);
    // row t of the N x N matrix of this plane (av1_qm_tables.h: 4x4 at 0, 8x8 at 16, 16x16 at 80)
    const uint8_t* qmw = (QM && P.qm[ss]) ? P.qm[ss] + (N == 4 ? 0 : N == 8 ? 16 : 80) + t * N : nullptr;
    // the dequantised row goes back to this thread's own row of buf (it has been consumed into row[])
#pragma unroll
    for (int l = 0; l < N; l++) {
      int64_t acc = 0;
#pragma unroll
      for (int j = 0; j < N / 2; j++) acc += (int64_t)TxTab<N>::f(l, j) * ((l & 1) ? rso[j] : rse[j]);
      const int32_t c = (int32_t)((acc * 4096 + ((int64_t)1 << (sh - 1))) >> sh);
      const bool dc = (t | l) == 0;
      uint32_t dqv = dc ? (uint32_t)P.dc_q : (uint32_t)P.ac_q;
      uint32_t lv;
      if (QM) {
        // the position's step: Round2(q * weight, 5); no quantisation matrix for a plane = weight 32 everywhere
        if (qmw) dqv = (dqv * (uint32_t)qmw[l] + 16u) >> 5;
        const uint32_t a = (uint32_t)(c < 0 ? -c : c) + ((dqv * (uint32_t)P.quant_rnd) >> 7);
        lv = a / dqv;
      } else {
        const uint32_t a = (uint32_t)(c < 0 ? -c : c) + (dc ? rnd_dc : rnd_ac);
        // exact floor(a / dqv) by multiplication with floor(2^32 / dqv) and at most two corrections
        lv = __umulhi(a, dc ? P.dc_magic : P.ac_magic);
        uint32_t rem = a - lv * dqv;
        if (rem >= dqv) { lv++; rem -= dqv; }
        if (rem >= dqv) lv++;
      }
      if (lv > 32767u) lv = 32767u;
      int32_t d = (int32_t)((lv * dqv) & 0xFFFFFFu);
      if (d > lim) d = lim;
      buf[t * S + l] = c < 0 ? -d : d;
      if (active) cdst[l] = (int16_t)(c < 0 ? -(int32_t)lv : (int32_t)lv);
      if (lv) eob = max(eob, TxTab<N>::iscan(t * N + l) + 1);
      sum_abs += (int)lv;
      lv8[t * S8 + l] = (uint8_t)min(lv, 15u);
      sign_bits |= (c < 0 ? 1u : 0u) << l;
      max_lv = max(max_lv, lv);
    }
    lv8[t * S8 + N] = 0; lv8[t * S8 + N + 1] = 0;
    if (t < 2) {
#pragma unroll
      for (int cidx = 0; cidx < S8; cidx++) lv8[(N + t) * S8 + cidx] = 0;
    }
#pragma unroll
    for (int o = N / 2; o; o >>= 1) {
      eob = max(eob, __shfl_xor_sync(gmask, eob, o));
      sum_abs += __shfl_xor_sync(gmask, sum_abs, o);
      max_lv = max(max_lv, __shfl_xor_sync(gmask, max_lv, o));
    }
    const int thr = N >= 16 ? P.tb_zero_thr : (N == 8 ? P.tb_zero_thr >> 1 : 0);
    if (eob > 0 && sum_abs <= thr) {
      // not worth coding: drop the whole transform block
      eob = 0;
      if (active) {
#pragma unroll
        for (int l = 0; l < N; l++) cdst[l] = 0;
      }
    }
  }
  // ---------------- pre-digested symbol stream for the host entropy coder ----------------
  // (levels up to 14 only: larger ones need Golomb escapes and stay in the raster form)
  // pack_levels == 2 (token path): every coded transform block, into the separate digest plane (levels saturate
  // at 15 there; the raster levels stay in place for the Golomb remainders)
  bool packed = false;
  const bool to_side = P.pack_levels == 2;
  if (P.pack_levels && eob > 0 && (to_side || max_lv < 15u)) {
    packed = !to_side;
    __syncwarp(gmask);
    if (active) {
      uint16_t* wdst = to_side ? P.digest[p] + (cbase - P.coef[p]) : reinterpret_cast<uint16_t*>(cbase);   // (cbase - coef includes fo)
#pragma unroll
      for (int l = 0; l < N; l++) {
        const int pos = t * N + l, si = TxTab<N>::iscan(pos);
        if (si < eob) {
          const uint8_t* L = lv8 + t * S8 + l;
          const int m3 = min((int)L[1], 3) + min((int)L[S8], 3) + min((int)L[S8 + 1], 3) + min((int)L[2], 3) + min((int)L[2 * S8], 3);
          const int bctx = pos == 0 ? 0 : min((m3 + 1) >> 1, 4) + TxTab<N>::nz_off(pos);
          const int m15 = (int)L[1] + (int)L[S8] + (int)L[S8 + 1];
          const int brctx = min((m15 + 1) >> 1, 6) + (pos == 0 ? 0 : ((t < 2 && l < 2) ? 7 : 14));
          wdst[si] = (uint16_t)((((sign_bits >> l) & 1u) << 15) | ((unsigned)L[0] << 11) | ((unsigned)brctx << 6) | (unsigned)bctx);
        }
      }
    }
  }
  // ---------------- reconstruction ----------------
  uint16_t* rec = P.rec[p] + fo + (size_t)y * stride + x;
  const int maxv = (1 << bd) - 1;
  if (eob == 0) {
    if (active) {
#pragma unroll
      for (int i = 0; i < N; i++) rec[(size_t)i * stride + t] = pred[i * N + t];
    }
    return 0;
  }
  const int eob_out = packed ? (eob | 0x8000) : eob;
  const int row_range = bd + 8, col_range = max(bd + 6, 16);
  {
    int32_t dq[N];
#pragma unroll
    for (int j = 0; j < N; j++) dq[j] = sat(buf[t * S + j], row_range);
    idct<N>(dq, row_range);
#pragma unroll
    for (int j = 0; j < N; j++) {
      int32_t v = dq[j];
      if (TxTab<N>::kRowShift > 0) v = (v + (1 << (TxTab<N>::kRowShift > 0 ? TxTab<N>::kRowShift - 1 : 0))) >> TxTab<N>::kRowShift;
      buf[t * S + j] = v;
    }
  }
  __syncwarp(gmask);
  {
    int32_t xc[N];
#pragma unroll
    for (int i = 0; i < N; i++) xc[i] = sat(buf[i * S + t], col_range);
    idct<N>(xc, col_range);
    if (active) {
#pragma unroll
      for (int i = 0; i < N; i++) {
        const int v = (xc[i] + 8) >> 4;
        rec[(size_t)i * stride + t] = (uint16_t)clampi((int)pred[i * N + t] + v, 0, maxv);
      }
    }
  }
  __syncwarp(gmask);
  return eob_out;
}

struct Smem {
  alignas(16) unsigned char pool[kPoolBytes];
  uint16_t eob[3][64];       // per plane, per 8x8 unit of the superblock (valid at a block's top-left unit)
  uint8_t bl[64];            // blk_log2 per unit, 0 = outside the picture
};

template <int N>
__device__ __forceinline__ void group_buffers(Smem& sm, int group, int32_t** buf, uint16_t** pred, uint8_t** lv8) {
  constexpr int kBuf = (N + 7) * (N + 1) * 4, kPred = N * N * 2, kLv = ((N + 2) * (N + 2) + 3) & ~3;
  unsigned char* base = sm.pool + (size_t)group * (kBuf + kPred + kLv);
  *buf = reinterpret_cast<int32_t*>(base);
  *pred = reinterpret_cast<uint16_t*>(base + kBuf);
  *lv8 = base + kBuf + kPred;
}

template <bool QM>
__global__ void __launch_bounds__(kThreads, 3) inter_encode_kernel(const InterLaunch P) {
  __shared__ Smem sm;
  // frame blockIdx.z of the launch: same reference, same quantiser, own source / outputs
  const int frame = blockIdx.z;
  const int16_t* mvs = P.mvs + (size_t)frame * P.map_elems * 2;
  const Av1bGeom& g = P.g;
  const int tid = threadIdx.x, lane = tid & 31;
  const int sbx = blockIdx.x, sby = blockIdx.y;
  if (tid < 64) {
    const int uy = sby * 8 + (tid >> 3), ux = sbx * 8 + (tid & 7);
    sm.bl[tid] = (uy < g.h8 && ux < g.w8) ? P.part_map[uy * g.w8 + ux] : 0;
    sm.eob[0][tid] = 0; sm.eob[1][tid] = 0; sm.eob[2][tid] = 0;
  }
  __syncthreads();
  // ---- luma of 16x16 blocks: 16 groups of 16 threads ----
  {
    const int group = tid >> 4, t = tid & 15;
    const unsigned gmask = 0xFFFFu << (lane & 16);
    const int u = (group >> 2) * 16 + (group & 3) * 2;      // top-left unit of the block
    const bool active = sm.bl[u] == 4;
    const int ux = sbx * 8 + (u & 7), uy = sby * 8 + (u >> 3);
    int mvr = 0, mvc = 0;
    if (active) { mvr = mvs[(uy * g.w8 + ux) * 2]; mvc = mvs[(uy * g.w8 + ux) * 2 + 1]; }
    int32_t* buf; uint16_t* pred;
    uint8_t* lv8;
    group_buffers<16>(sm, group, &buf, &pred, &lv8);
    // inactive groups run on the (always allocated) superblock origin and store nothing
    const int x = active ? ux * 8 : sbx * 64, y = active ? uy * 8 : sby * 64;
    const int eob = code_tb<16, QM>(P, frame, 0, x, y, mvr, mvc, t, gmask, buf, pred, lv8, active);
    if (active && t == 0) sm.eob[0][u] = (uint16_t)eob;
  }
  __syncthreads();
  // ---- chroma of 16x16 blocks: 32 groups of 8 threads (16 blocks x U, V) ----
  {
    const int group = tid >> 3, t = tid & 7;
    const unsigned gmask = 0xFFu << (lane & 24);
    const int b = group & 15, p = 1 + (group >> 4);
    const int u = (b >> 2) * 16 + (b & 3) * 2;
    const bool active = sm.bl[u] == 4;
    const int ux = sbx * 8 + (u & 7), uy = sby * 8 + (u >> 3);
    int mvr = 0, mvc = 0;
    if (active) { mvr = mvs[(uy * g.w8 + ux) * 2]; mvc = mvs[(uy * g.w8 + ux) * 2 + 1]; }
    int32_t* buf; uint16_t* pred;
    uint8_t* lv8;
      group_buffers<8>(sm, group, &buf, &pred, &lv8);
    const int x = active ? ux * 4 : sbx * 32, y = active ? uy * 4 : sby * 32;
    const int eob = code_tb<8, QM>(P, frame, p, x, y, mvr, mvc, t, gmask, buf, pred, lv8, active);
    if (active && t == 0) sm.eob[p][u] = (uint16_t)eob;
  }
  __syncthreads();
  // ---- 8x8 blocks (picture edge): luma 8x8 in groups of 8 threads, two rounds of 32 units ----
  bool any8 = false;
  for (int i = lane; i < 64; i += 32) any8 |= sm.bl[i] == 3;
  any8 = __any_sync(0xffffffffu, any8);
  if (any8) {
    for (int round = 0; round < 2; round++) {
      const int group = tid >> 3, t = tid & 7;
      const unsigned gmask = 0xFFu << (lane & 24);
      const int u = round * 32 + group;
      const bool active = sm.bl[u] == 3;
      const int ux = sbx * 8 + (u & 7), uy = sby * 8 + (u >> 3);
      int mvr = 0, mvc = 0;
      if (active) { mvr = mvs[(uy * g.w8 + ux) * 2]; mvc = mvs[(uy * g.w8 + ux) * 2 + 1]; }
      int32_t* buf; uint16_t* pred;
      uint8_t* lv8;
      group_buffers<8>(sm, group, &buf, &pred, &lv8);
      const int x = active ? ux * 8 : sbx * 64, y = active ? uy * 8 : sby * 64;
      const int eob = code_tb<8, QM>(P, frame, 0, x, y, mvr, mvc, t, gmask, buf, pred, lv8, active);
      if (active && t == 0) sm.eob[0][u] = (uint16_t)eob;
      __syncthreads();
    }
    // chroma 4x4 of the 8x8 blocks: 64 groups of 4 threads, one round per plane
    for (int p = 1; p < 3; p++) {
      const int group = tid >> 2, t = tid & 3;
      const unsigned gmask = 0xFu << (lane & 28);
      const int u = group;
      const bool active = sm.bl[u] == 3;
      const int ux = sbx * 8 + (u & 7), uy = sby * 8 + (u >> 3);
      int mvr = 0, mvc = 0;
      if (active) { mvr = mvs[(uy * g.w8 + ux) * 2]; mvc = mvs[(uy * g.w8 + ux) * 2 + 1]; }
      int32_t* buf; uint16_t* pred;
      uint8_t* lv8;
      group_buffers<4>(sm, group, &buf, &pred, &lv8);
      const int x = active ? ux * 4 : sbx * 32, y = active ? uy * 4 : sby * 32;
      const int eob = code_tb<4, QM>(P, frame, p, x, y, mvr, mvc, t, gmask, buf, pred, lv8, active);
      if (active && t == 0) sm.eob[p][u] = (uint16_t)eob;
      __syncthreads();
    }
  }
  // ---- block side information ----
  if (tid < 64) {
    const int u = tid, bl = sm.bl[u];
    if (bl) {
      const int n8 = 1 << (bl - 3);
      const int ox = (u & 7) & ~(n8 - 1), oy = (u >> 3) & ~(n8 - 1), o = oy * 8 + ox;   // block's top-left unit
      const int ux = sbx * 8 + ox, uy = sby * 8 + oy;
      Av1bBlockInfo info;
      info.blk_log2 = (uint8_t)bl; info.y_mode = 0; info.uv_mode = 0;
      info.eob[0] = sm.eob[0][o]; info.eob[1] = sm.eob[1][o]; info.eob[2] = sm.eob[2][o];
      info.skip = (info.eob[0] | info.eob[1] | info.eob[2]) == 0;
      info.angle_y = 0; info.angle_uv = 0; info.tx_type_y = AV1B_DCT_DCT; info.cfl_alpha_u = 0; info.cfl_alpha_v = 0;
      info.is_inter = 1;
      info.mv[0] = mvs[(uy * g.w8 + ux) * 2]; info.mv[1] = mvs[(uy * g.w8 + ux) * 2 + 1];
      P.blocks[(size_t)frame * P.map_elems + (sby * 8 + (u >> 3)) * g.w8 + sbx * 8 + (u & 7)] = info;
    }
  }
}

}  // namespace

// One CTA of 64 threads per superblock: thread u owns 8x8 unit u (row-major inside the superblock).
__global__ void __launch_bounds__(64) merge_skip_kernel(Av1bGeom g, Av1bBlockInfo* blocks0, size_t map_elems) {
  Av1bBlockInfo* blocks = blocks0 + (size_t)blockIdx.z * map_elems;
  __shared__ uint8_t bl[64], ok16[64];
  __shared__ int16_t mv[64][2];
  const int u = threadIdx.x, ux = blockIdx.x * 8 + (u & 7), uy = blockIdx.y * 8 + (u >> 3);
  const bool inside = ux < g.w8 && uy < g.h8;
  Av1bBlockInfo* b = inside ? blocks + (size_t)uy * g.w8 + ux : nullptr;
  bl[u] = inside ? b->blk_log2 : 0;
  ok16[u] = inside && b->is_inter && b->skip;
  mv[u][0] = inside ? b->mv[0] : 0; mv[u][1] = inside ? b->mv[1] : 0;
  __syncthreads();
  if (u < 4) {
    // 32x32 quadrant u: children are the 16x16 blocks whose top-left units are o, o+2, o+16, o+18
    const int o = (u >> 1) * 32 + (u & 1) * 4;
    bool ok = true;
    for (int q = 0; q < 4; q++) {
      const int c = o + (q >> 1) * 16 + (q & 1) * 2;
      ok = ok && bl[c] == 4 && ok16[c] && mv[c][0] == mv[o][0] && mv[c][1] == mv[o][1];
    }
    // the whole 32x32 must lie inside the picture: its bottom-right unit exists
    ok = ok && bl[o + 27] != 0;
    if (ok) for (int yy = 0; yy < 4; yy++) for (int xx = 0; xx < 4; xx++) bl[o + yy * 8 + xx] = 5;
  }
  __syncthreads();
  if (u == 0) {
    bool ok = bl[63] != 0;
    for (int q = 0; q < 4; q++) {
      const int c = (q >> 1) * 32 + (q & 1) * 4;
      ok = ok && bl[c] == 5 && mv[c][0] == mv[0][0] && mv[c][1] == mv[0][1];
    }
    if (ok) for (int i = 0; i < 64; i++) bl[i] = 6;
  }
  __syncthreads();
  if (inside && b->blk_log2 != bl[u]) b->blk_log2 = bl[u];
}

cudaError_t launch_merge_skip(const Av1bGeom& g, Av1bBlockInfo* blocks, size_t map_elems, int n_frames, cudaStream_t s) {
  dim3 grid(g.sb_cols, g.sb_rows, n_frames);
  merge_skip_kernel<<<grid, 64, 0, s>>>(g, blocks, map_elems);
  return cudaGetLastError();
}

cudaError_t launch_inter_encode(const InterLaunch& p0, cudaStream_t s) {
  InterLaunch p = p0;
  p.dc_magic = (uint32_t)(((uint64_t)1 << 32) / (uint32_t)p.dc_q);
  p.ac_magic = (uint32_t)(((uint64_t)1 << 32) / (uint32_t)p.ac_q);
  dim3 grid(p.g.sb_cols, p.g.sb_rows, p.n_frames > 0 ? p.n_frames : 1);
  if (p.qm[0] || p.qm[1]) inter_encode_kernel<true><<<grid, kThreads, 0, s>>>(p);
  else inter_encode_kernel<false><<<grid, kThreads, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace av1b
