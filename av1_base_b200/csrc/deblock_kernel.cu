// AV1 deblocking loop filter (spec 7.14) for sm_100a: one CTA per 64x64 superblock per frame.
// The CTA stages the superblock plus an 8-sample apron of each plane in shared memory, filters all
// vertical edges (every row of the staged window), then all horizontal edges (the tile's own
// columns), and writes only its own 64x64 (32x32 chroma) samples to the OUTPUT frame.  Because AV1
// limits every edge filter to min(tx size on both sides) no two edge filters of one pass touch the
// same sample, and a filter never reads a sample another filter of the same pass writes; the window
// therefore reproduces the whole-frame two-pass definition exactly while the frame is read once and
// written once (algorithmic bytes 2*S).
//
// Replaces arithmetic the reference delegates to av1an + SVT-AV1
// (/root/reference/crates/daemon/src/encode/av1an.rs:126-139).
// Bit-exact against oracle/av1_oracle.cpp orc_deblock_frame (pinned vs libaom aom_highbd_lpf_*_c
// and both decoders).
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.cuh"

namespace av1b {
namespace {

constexpr int kThreads = 256;
constexpr int kApron = 8;
constexpr int kMaxWin = 64 + 2 * kApron;        // 80
constexpr int kWinStride = kMaxWin + 2;         // 82 samples = 41 words: conflict-free column walks

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int iabs(int v) { return v < 0 ? -v : v; }

struct LfThresh { int limit, blimit, thresh; };

__device__ __forceinline__ LfThresh lf_thresh(int lvl, int sharp, int bd) {
  const int shift = sharp > 4 ? 2 : (sharp > 0 ? 1 : 0);
  int limit = sharp > 0 ? clampi(lvl >> shift, 1, 9 - sharp) : max(1, lvl >> shift);
  LfThresh t;
  t.limit = limit << (bd - 8);
  t.blimit = (2 * (lvl + 2) + limit) << (bd - 8);
  t.thresh = (lvl >> 4) << (bd - 8);
  return t;
}

// wide filter: window of 2n+1 taps, the centre 2*n2+1 taps doubled, taps clamped to [-(n+1), n]
template <int N, int N2, int LOG2>
__device__ __forceinline__ void lf_wide(uint16_t* s, int step) {
  int in[2 * N + 2];
#pragma unroll
  for (int i = 0; i < 2 * N + 2; i++) in[i] = s[(i - (N + 1)) * step];
  int out[2 * N];
#pragma unroll
  for (int i = -N; i < N; i++) {
    int t = 0;
#pragma unroll
    for (int j = -N; j <= N; j++) {
      int p = i + j;
      p = p < -(N + 1) ? -(N + 1) : (p > N ? N : p);
      t += in[p + N + 1] * ((j >= -N2 && j <= N2) ? 2 : 1);
    }
    out[i + N] = (t + (1 << (LOG2 - 1))) >> LOG2;
  }
#pragma unroll
  for (int i = -N; i < N; i++) s[i * step] = (uint16_t)out[i + N];
}

// one sample line across an edge: s points at q0, s[-step] is p0.  fsz: 4, 6 (chroma), 8 or 16.
__device__ __forceinline__ void lf_line(uint16_t* s, int step, int fsz, const LfThresh& T, int bd) {
  const int one = 1 << (bd - 8);
  const int p0 = s[-step], p1 = s[-2 * step], q0 = s[0], q1 = s[step];
  const bool hev = iabs(p1 - p0) > T.thresh || iabs(q1 - q0) > T.thresh;
  bool mask = iabs(p1 - p0) <= T.limit && iabs(q1 - q0) <= T.limit &&
              iabs(p0 - q0) * 2 + iabs(p1 - q1) / 2 <= T.blimit;
  bool flat = false, flat2 = false;
  if (fsz >= 6) {
    const int p2 = s[-3 * step], q2 = s[2 * step];
    mask = mask && iabs(p2 - p1) <= T.limit && iabs(q2 - q1) <= T.limit;
    flat = iabs(p1 - p0) <= one && iabs(q1 - q0) <= one && iabs(p2 - p0) <= one && iabs(q2 - q0) <= one;
    if (fsz >= 8) {
      const int p3 = s[-4 * step], q3 = s[3 * step];
      mask = mask && iabs(p3 - p2) <= T.limit && iabs(q3 - q2) <= T.limit;
      flat = flat && iabs(p3 - p0) <= one && iabs(q3 - q0) <= one;
    }
  }
  if (!mask) return;
  if (fsz == 16 && flat) {
    flat2 = iabs((int)s[-5 * step] - p0) <= one && iabs((int)s[-6 * step] - p0) <= one &&
            iabs((int)s[-7 * step] - p0) <= one && iabs((int)s[4 * step] - q0) <= one &&
            iabs((int)s[5 * step] - q0) <= one && iabs((int)s[6 * step] - q0) <= one;
  }
  if (fsz == 4 || !flat) {
    const int lo = -(1 << (bd - 1)), hi = (1 << (bd - 1)) - 1, off = 0x80 << (bd - 8);
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
  if (fsz == 16 && flat2) lf_wide<6, 1, 4>(s, step);
  else if (fsz == 6) lf_wide<2, 1, 3>(s, step);
  else lf_wide<3, 0, 3>(s, step);
}

struct Smem {
  uint16_t win[kMaxWin * kWinStride];
  uint8_t bl[10][10];     // blk_log2 of the 8x8 units around the superblock ([uy+1][ux+1]); 0 = outside
};

__global__ void __launch_bounds__(kThreads) deblock_kernel(const DeblockLaunch P) {
  __shared__ Smem sm;
  const Av1bGeom& g = P.g;
  const int tid = threadIdx.x;
  const int sbx = blockIdx.x, sby = blockIdx.y, frame = blockIdx.z;
  const Av1bBlockInfo* blocks = P.blocks + (size_t)frame * P.map_elems;
  const int bd = P.bit_depth;
  if (tid < 100) {
    const int uy = sby * 8 + tid / 10 - 1, ux = sbx * 8 + tid % 10 - 1;
    uint8_t v = 0;
    if (uy >= 0 && ux >= 0 && uy < g.h8 && ux < g.w8) v = blocks[uy * g.w8 + ux].blk_log2;
    sm.bl[tid / 10][tid % 10] = v;
  }
  for (int p = 0; p < 3; p++) {
    const int ss = p > 0;
    const bool lf_off = !P.lf_level[0] && !P.lf_level[1];   // loop filter disabled for the whole frame
    const int lvl_v = lf_off ? 0 : (p == 0 ? P.lf_level[0] : P.lf_level[p + 1]);
    const int lvl_h = lf_off ? 0 : (p == 0 ? P.lf_level[1] : P.lf_level[p + 1]);
    const int T = 64 >> ss, W = T + 2 * kApron;
    const int stride = g.stride[p], rows = g.rows[p];
    const int pw = (g.mi_cols * 4) >> ss, ph = (g.mi_rows * 4) >> ss;
    const int x0 = sbx * T, y0 = sby * T;
    const uint16_t* in = P.in[p] + (size_t)frame * P.plane_elems[p];
    uint16_t* out = P.out[p] + (size_t)frame * P.plane_elems[p];
    __syncthreads();   // previous plane's window fully consumed; bl[] visible
    if (!lvl_v && !lvl_h) {
      // plane not filtered: plain copy of the tile
      for (int o = tid; o < T * (T / 8); o += kThreads) {
        const int r = o / (T / 8), v = o % (T / 8);
        const size_t off = (size_t)(y0 + r) * stride + x0 + v * 8;
        *reinterpret_cast<uint4*>(out + off) = *reinterpret_cast<const uint4*>(in + off);
      }
      continue;
    }
    // ---- stage the window: rows y0-8 .. y0+T+7, cols x0-8 .. x0+T+7 (16-byte vectors) ----
    const int vecs = W / 8;
    for (int o = tid; o < W * vecs; o += kThreads) {
      const int r = o / vecs, v = o % vecs;
      const int y = y0 - kApron + r, x = x0 - kApron + v * 8;
      uint4 d = make_uint4(0, 0, 0, 0);
      if (y >= 0 && y < rows && x >= 0 && x < stride) d = *reinterpret_cast<const uint4*>(in + (size_t)y * stride + x);
      uint32_t* w = reinterpret_cast<uint32_t*>(sm.win + r * kWinStride + v * 8);
      w[0] = d.x; w[1] = d.y; w[2] = d.z; w[3] = d.w;
    }
    __syncthreads();
    const int unit = 8 >> ss;                 // samples of this plane per 8x8 luma unit
    const int lu = 3 - ss;                    // log2(unit)
    const int nslots = T / unit + 1;          // candidate edge positions: x0, x0+unit, ..., x0+T
    const int txmax = 64 >> ss, fmax = ss ? 8 : 16;
    // ---- vertical edges: every staged row ----
    if (lvl_v) {
      const LfThresh th = lf_thresh(lvl_v, P.sharpness, bd);
      for (int o = tid; o < W * nslots; o += kThreads) {
        const int r = o % W, e = o / W;       // consecutive lanes walk down a column of the window
        const int y = y0 - kApron + r, xe = x0 + e * unit;
        // rows further than one 8x8 unit from the tile are never read by the horizontal pass
        if (y < y0 - unit || y >= y0 + T + unit || y < 0 || y >= ph || xe <= 0 || xe >= pw) continue;
        const int uy = (y >> lu) - sby * 8 + 1, ux = (xe >> lu) - sbx * 8 + 1;
        const int bc = sm.bl[uy][ux], bp = sm.bl[uy][ux - 1];
        const int txc = min(1 << (bc - ss), txmax), txp = min(1 << (bp - ss), txmax);
        if (xe & (txc - 1)) continue;
        const int fs = min(min(txc, txp), fmax);
        const int fsz = ss ? (fs == 8 ? 6 : 4) : fs;
        lf_line(sm.win + r * kWinStride + kApron + e * unit, 1, fsz, th, bd);
      }
    }
    __syncthreads();
    // ---- horizontal edges: the tile's own columns ----
    if (lvl_h) {
      const LfThresh th = lf_thresh(lvl_h, P.sharpness, bd);
      for (int o = tid; o < T * nslots; o += kThreads) {
        const int c = o % T, e = o / T;
        const int x = x0 + c, ye = y0 + e * unit;
        if (x >= pw || ye <= 0 || ye >= ph) continue;
        const int uy = (ye >> lu) - sby * 8 + 1, ux = (x >> lu) - sbx * 8 + 1;
        const int bc = sm.bl[uy][ux], bp = sm.bl[uy - 1][ux];
        const int txc = min(1 << (bc - ss), txmax), txp = min(1 << (bp - ss), txmax);
        if (ye & (txc - 1)) continue;
        const int fs = min(min(txc, txp), fmax);
        const int fsz = ss ? (fs == 8 ? 6 : 4) : fs;
        lf_line(sm.win + (kApron + e * unit) * kWinStride + kApron + c, kWinStride, fsz, th, bd);
      }
    }
    __syncthreads();
    // ---- write the tile ----
    for (int o = tid; o < T * (T / 8); o += kThreads) {
      const int r = o / (T / 8), v = o % (T / 8);
      const uint32_t* w = reinterpret_cast<const uint32_t*>(sm.win + (kApron + r) * kWinStride + kApron + v * 8);
      *reinterpret_cast<uint4*>(out + (size_t)(y0 + r) * stride + x0 + v * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

}  // namespace

cudaError_t launch_deblock(const DeblockLaunch& p, int n_frames, cudaStream_t s) {
  dim3 grid(p.g.sb_cols, p.g.sb_rows, n_frames);
  deblock_kernel<<<grid, kThreads, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace av1b
