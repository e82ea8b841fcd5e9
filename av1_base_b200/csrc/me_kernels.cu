// Source pyramid and hierarchical motion estimation (open loop, on SOURCE pictures) for sm_100a.
//
//   pyramid_kernel : L0 luma -> L1 (1/2) and L2 (1/4) in one pass; one thread per L2 sample reads a
//                    4x4 patch (four 8-byte loads), writes 2x2 L1 samples and one L2 sample.
//                    Streaming, HBM-bound: algorithmic bytes 1.3125 * Y.
//   hme_l2_kernel  : full search +-12 on the 1/4 picture for every 8x8 block (= 32x32 luma), on 8-bit samples.  One
//                    CTA stages a 32x32 patch of the current picture and its (32+24)^2 search window of the
//                    reference in shared memory as bytes; one warp per block, lane = candidate column, the 25
//                    candidate rows of a lane accumulate in registers from a sliding window row (VABSDIFF4.U8.ACC),
//                    warp arg-min on (cost, visiting order).
//   hme_refine_kernel : per 16x16 luma block, +-2 on L1 around twice the L2 vector, then +-2 on L0
//                    around twice the L1 vector; one warp per block, windows staged in shared memory.
//
// Replaces arithmetic the reference delegates to av1an + SVT-AV1
// (/root/reference/crates/daemon/src/encode/av1an.rs:126-139).
// Decisions are defined by oracle/av1_oracle.cpp (orc_downscale2, orc_hme) and match it bit for bit.
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.cuh"

namespace av1b {
namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pyramid_kernel(const uint16_t* __restrict__ l0, uint16_t* __restrict__ l1,
                                                      uint16_t* __restrict__ l2, int stride0, int rows0,
                                                      size_t elems0, int n_frames) {
  const int s2 = stride0 >> 2, r2 = rows0 >> 2;
  const size_t per = (size_t)s2 * r2;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per * n_frames) return;
  const int f = (int)(idx / per);
  const int rem = (int)(idx - (size_t)f * per), y2 = rem / s2, x2 = rem - y2 * s2;
  const uint16_t* src = l0 + (size_t)f * elems0 + (size_t)(4 * y2) * stride0 + 4 * x2;
  uint32_t a[4][2];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uint2 v = *reinterpret_cast<const uint2*>(src + (size_t)r * stride0);
    a[r][0] = v.x; a[r][1] = v.y;
  }
  uint32_t q[2][2];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const uint32_t t = a[2 * i][j], b = a[2 * i + 1][j];
      q[i][j] = ((t & 0xFFFF) + (t >> 16) + (b & 0xFFFF) + (b >> 16) + 2) >> 2;
    }
  uint16_t* d1 = l1 + (size_t)f * (elems0 >> 2) + (size_t)(2 * y2) * (stride0 >> 1) + 2 * x2;
  *reinterpret_cast<uint32_t*>(d1) = q[0][0] | (q[0][1] << 16);
  *reinterpret_cast<uint32_t*>(d1 + (stride0 >> 1)) = q[1][0] | (q[1][1] << 16);
  l2[(size_t)f * (elems0 >> 4) + (size_t)y2 * s2 + x2] = (uint16_t)((q[0][0] + q[0][1] + q[1][0] + q[1][1] + 2) >> 2);
}

// ---------------------------------------------------------------------------------------------
constexpr int kR2 = 12;
constexpr int kT2 = 32;                     // L2 samples per CTA side (4x4 blocks of 8x8)
constexpr int kW2 = kT2 + 2 * kR2;          // 56
constexpr int kW2P = 64;                    // staged window row pitch in bytes (16 words)

// The quarter-resolution search works on 8-bit samples, min(v >> shift2, 255) (shift2 = bit depth - 8): four samples
// per 32-bit word and one VABSDIFF4.U8.ACC per four absolute differences.
struct L2Smem {
  uint32_t cur[kT2 * kT2 / 4];
  uint32_t ref[kW2 * kW2P / 4];
};

__global__ void __launch_bounds__(256) hme_l2_kernel(const HmeLaunch P) {
  __shared__ L2Smem sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int frame = blockIdx.z;
  const int w2 = P.width >> 2, h2 = P.height >> 2, s2 = P.stride0 >> 2;
  const size_t e2 = P.elems0 >> 4;
  const uint16_t* cur = P.cur[2] + (size_t)P.cur_slot[frame] * e2;
  const uint16_t* ref = P.ref[2] + (size_t)P.ref_slot[frame] * e2;
  const int x0 = blockIdx.x * kT2, y0 = blockIdx.y * kT2;
  const int sh = P.shift2;
  // one thread packs four neighbouring samples into a word
  for (int o = tid; o < kT2 * kT2 / 4; o += 256) {
    const int r = o / (kT2 / 4), c = (o % (kT2 / 4)) * 4;
    const uint16_t* row = cur + (size_t)clampi(y0 + r, 0, h2 - 1) * s2;
    uint32_t w = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) w |= min((uint32_t)row[clampi(x0 + c + j, 0, w2 - 1)] >> sh, 255u) << (8 * j);
    sm.cur[o] = w;
  }
  for (int o = tid; o < kW2 * (kW2 / 4); o += 256) {
    const int r = o / (kW2 / 4), c = (o % (kW2 / 4)) * 4;
    const uint16_t* row = ref + (size_t)clampi(y0 - kR2 + r, 0, h2 - 1) * s2;
    uint32_t w = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) w |= min((uint32_t)row[clampi(x0 - kR2 + c + j, 0, w2 - 1)] >> sh, 255u) << (8 * j);
    sm.ref[r * (kW2P / 4) + (c >> 2)] = w;
  }
  __syncthreads();
  const int n2x = (P.width + 31) / 32, n2y = (P.height + 31) / 32;
  constexpr int kSide = 2 * kR2 + 1, kCand = kSide * kSide, kCentre = kR2 * kSide + kR2;
  const int lam2 = max(1, (P.lambda >> sh) >> 4);
  for (int b = warp; b < 16; b += 8) {
    const int bx = b & 3, by = b >> 2;
    const int gbx = blockIdx.x * 4 + bx, gby = blockIdx.y * 4 + by;
    if (gbx >= n2x || gby >= n2y) continue;
    // current block in registers: 8 rows of two words
    uint32_t c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = sm.cur[(by * 8 + i) * (kT2 / 4) + bx * 2]; c[i][1] = sm.cur[(by * 8 + i) * (kT2 / 4) + bx * 2 + 1]; }
    unsigned best = 0xFFFFFFFFu;   // (cost << 10) | visiting order
    // lane = candidate column (dx + 12), all 25 candidate rows in registers: a window row is read once (three aligned
    // words, shifted to the lane's byte offset) and feeds the up to eight candidate rows that overlap it
    if (lane < kSide) {
      unsigned acc[kSide];
#pragma unroll
      for (int k = 0; k < kSide; k++) acc[k] = 0;
      const int boff = bx * 8 + lane;                 // first byte of the lane's 8-sample row inside the window row
      const uint32_t* rp0 = sm.ref + (by * 8) * (kW2P / 4) + (boff >> 2);
      const int fs = (boff & 3) * 8;
#pragma unroll
      for (int wr = 0; wr < 8 + 2 * kR2; wr++) {
        const uint32_t w0 = rp0[wr * (kW2P / 4)], w1 = rp0[wr * (kW2P / 4) + 1], w2w = rp0[wr * (kW2P / 4) + 2];
        const uint32_t s0 = __funnelshift_r(w0, w1, fs), s1 = __funnelshift_r(w1, w2w, fs);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int dyi = wr - i;
          if (dyi >= 0 && dyi < kSide) acc[dyi] = __vsadu4(c[i][1], s1) + (__vsadu4(c[i][0], s0) + acc[dyi]);
        }
      }
      const int dx = lane - kR2;
#pragma unroll
      for (int dyi = 0; dyi < kSide; dyi++) {
        const int k = dyi * kSide + lane, dy = dyi - kR2;
        // visiting order: centre first, then raster
        const int order = k == kCentre ? 0 : (k < kCentre ? k + 1 : k);
        const int cost = (int)acc[dyi] + lam2 * (abs(dx) + abs(dy));
        best = min(best, ((unsigned)cost << 10) | (unsigned)order);
      }
    }
    best = __reduce_min_sync(0xffffffffu, best);
    if (lane == 0) {
      const int order = best & 1023;
      const int k = order == 0 ? kCentre : (order <= kCentre ? order - 1 : order);
      int16_t* out = P.mv2 + ((size_t)frame * n2x * n2y + (size_t)gby * n2x + gbx) * 2;
      out[0] = (int16_t)(k / kSide - kR2);
      out[1] = (int16_t)(k % kSide - kR2);
    }
  }
}

// ---------------------------------------------------------------------------------------------
struct RefineSmem {
  alignas(16) uint16_t cur1[8][8 * 8];
  alignas(4) uint16_t ref1[8][12 * 14];   // row pitch 14 (7 words): 12 samples + an odd start
  alignas(16) uint16_t cur0[8][16 * 16];
  alignas(4) uint16_t ref0[8][20 * 24];   // row pitch 24 (12 words): 20 samples + an odd start (sad25_rows16 reads words)
};

// 25 candidates (+-2), lanes = candidates; returns the chosen (dy, dx) in all lanes
__device__ __forceinline__ int subpel_parabola(int sm, int s0, int sp, int lambda) {
  const int num = sm - sp, den = 2 * (sm - 2 * s0 + sp);
  if (den <= 0) return 0;
  if ((long long)num * num <= (long long)4 * den * lambda) return 0;
  const int a = 8 * num + den, b = 2 * den;
  const int q = a >= 0 ? a / b : -((-a + b - 1) / b);
  return clampi(q, -2, 2);
}

// The 25 SADs of a 16x16 block with the lanes laid over the BLOCK instead of the candidates: lane (r, h) owns
// row r, columns 8h .. 8h+7 of the current block in registers, reads the five window rows r .. r+4 as aligned
// 32-bit words (12 samples each) and accumulates its 8-sample share of all 25 candidates; a reduce-scatter adds the
// shares.  39 shared-memory loads per lane instead of 512 (the candidate-per-lane form was bound by them).
// Window row pitch: 22 samples (rows start word aligned).  Lane k < 25 returns the SAD of candidate k.
constexpr int kRS0 = 24;
__device__ __forceinline__ int sad25_rows16(const uint16_t* cur, const uint16_t* ref, int lane, int xo) {
  const int r = lane >> 1, h = lane & 1;
  unsigned c[8];
  {
    const uint32_t* cw = reinterpret_cast<const uint32_t*>(cur + r * 16 + 8 * h);
#pragma unroll
    for (int k = 0; k < 4; k++) { const uint32_t w = cw[k]; c[2 * k] = w & 0xFFFFu; c[2 * k + 1] = w >> 16; }
  }
  unsigned acc[25];
#pragma unroll
  for (int k = 0; k < 25; k++) acc[k] = 0;
#pragma unroll
  for (int wr = 0; wr < 5; wr++) {
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(ref + (r + wr) * kRS0 + 8 * h);
    uint32_t w[7];
#pragma unroll
    for (int k = 0; k < 6; k++) w[k] = rw[k];
    if (xo) {   // warp-uniform: the window starts one sample into its first word
      w[6] = rw[6];
#pragma unroll
      for (int k = 0; k < 6; k++) w[k] = __funnelshift_r(w[k], w[k + 1], 16);
    }
    unsigned s[12];
#pragma unroll
    for (int k = 0; k < 6; k++) { s[2 * k] = w[k] & 0xFFFFu; s[2 * k + 1] = w[k] >> 16; }
#pragma unroll
    for (int dx = 0; dx < 5; dx++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[wr * 5 + dx] = __usad(c[j], s[dx + j], acc[wr * 5 + dx]);
  }
  // reduce-scatter over the lanes: at step o the lanes with bit o set keep the upper half of the (padded to 32)
  // candidate list and hand the lower half to their partner, and vice versa; after five steps lane k holds the
  // total of candidate k (31 exchanges instead of 125 for a full butterfly of every candidate)
  unsigned v[32];
#pragma unroll
  for (int k = 0; k < 32; k++) v[k] = k < 25 ? acc[k] : 0u;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const bool hi = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; i++) {
      const unsigned a = v[i], b = v[i + o];
      const unsigned keep = hi ? b : a, send = hi ? a : b;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return (int)v[0];
}

template <int N, int RS, bool kSubpel>
__device__ __forceinline__ void refine25(const uint16_t* cur, const uint16_t* ref, int xo, int lane, int lam, int lam_sub,
                                         int* bdy, int* bdx, int* qy, int* qx) {
  unsigned key = 0xFFFFFFFFu;
  int my_sad = 0;
  int pre_sad = 0;
  if (N == 16) pre_sad = sad25_rows16(cur, ref, lane, xo);   // all lanes take part
  if (lane < 25) {
    const int dy = lane / 5 - 2, dx = lane % 5 - 2;
    unsigned usad = 0;
    if (N == 16) {
      usad = (unsigned)pre_sad;
    } else {
      const uint16_t* rp = ref + (2 + dy) * RS + 2 + dx + xo;
      for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) usad = __usad((unsigned)cur[i * N + j], (unsigned)rp[i * RS + j], usad);
    }
    const int sad = (int)usad;
    const int order = lane == 12 ? 0 : (lane < 12 ? lane + 1 : lane);
    key = ((unsigned)(sad + lam * (abs(dy) + abs(dx))) << 5) | (unsigned)order;
    my_sad = sad;
  }
  for (int o = 16; o; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
  const int order = key & 31;
  const int k = order == 0 ? 12 : (order <= 12 ? order - 1 : order);
  *bdy = k / 5 - 2;
  *bdx = k % 5 - 2;
  if (kSubpel) {
    // quarter-sample offset per axis from the parabola through the SADs next to the winner
    const int s0 = __shfl_sync(0xffffffffu, my_sad, k);
    const int sl = __shfl_sync(0xffffffffu, my_sad, (k + 31) & 31), sr = __shfl_sync(0xffffffffu, my_sad, (k + 1) & 31);
    const int su = __shfl_sync(0xffffffffu, my_sad, (k + 27) & 31), sd = __shfl_sync(0xffffffffu, my_sad, (k + 5) & 31);
    *qx = (k % 5 >= 1 && k % 5 <= 3) ? subpel_parabola(sl, s0, sr, lam_sub) : 0;
    *qy = (k / 5 >= 1 && k / 5 <= 3) ? subpel_parabola(su, s0, sd, lam_sub) : 0;
  }
}

__global__ void __launch_bounds__(256) hme_refine_kernel(const HmeLaunch P) {
  __shared__ RefineSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int frame = blockIdx.z;
  const int n1x = (P.width + 15) / 16, n1y = (P.height + 15) / 16;
  const int n2x = (P.width + 31) / 32, n2y = (P.height + 31) / 32;
  const int blk = blockIdx.x * 8 + warp;
  if (blk >= n1x * n1y) return;
  const int bx = blk % n1x, by = blk / n1x;
  const int w1 = P.width >> 1, h1 = P.height >> 1, s1 = P.stride0 >> 1;
  const uint16_t* cur1 = P.cur[1] + (size_t)P.cur_slot[frame] * (P.elems0 >> 2);
  const uint16_t* ref1 = P.ref[1] + (size_t)P.ref_slot[frame] * (P.elems0 >> 2);
  const uint16_t* cur0 = P.cur[0] + (size_t)P.cur_slot[frame] * P.elems0;
  const uint16_t* ref0 = P.ref[0] + (size_t)P.ref_slot[frame] * P.elems0;
  const int16_t* m2 = P.mv2 + ((size_t)frame * n2x * n2y + (size_t)(by >> 1) * n2x + (bx >> 1)) * 2;
  const int py = 2 * m2[0], px = 2 * m2[1];
  // ---- L1: 8x8 block at (8bx, 8by), +-2 around (px, py) ----
  // windows that lie inside the picture (almost all) are staged without clamping, the block itself as 16-byte rows
  const bool in1 = by * 8 + 8 <= h1 && bx * 8 + 8 <= w1 && by * 8 + py - 2 >= 0 && by * 8 + py + 10 <= h1 &&
                   bx * 8 + px - 2 >= 0 && bx * 8 + px + 10 <= w1 && bx * 8 + px + 12 <= s1;
  int xo1 = 0;
  if (in1) {
    if (lane < 8)
      *reinterpret_cast<uint4*>(&sm.cur1[warp][lane * 8]) = *reinterpret_cast<const uint4*>(cur1 + (size_t)(by * 8 + lane) * s1 + bx * 8);
    // 12 rows of 12 samples as aligned words (7 per row cover an odd start): lane = (row of four, word)
    const int wx = bx * 8 + px - 2;
    xo1 = wx & 1;
    const int wr = lane >> 3, wc = lane & 7;
    if (wc < 7) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(ref1 + (size_t)(by * 8 + py - 2 + wr) * s1 + (wx - xo1)) + wc;
      uint32_t* dst = reinterpret_cast<uint32_t*>(sm.ref1[warp]) + wr * 7 + wc;
#pragma unroll
      for (int i = 0; i < 3; i++) dst[i * 28] = src[(size_t)i * 2 * s1];   // rows wr, wr + 4, wr + 8 (s1 samples = s1 / 2 words)
    }
  } else {
    for (int o = lane; o < 64; o += 32)
      sm.cur1[warp][o] = cur1[(size_t)clampi(by * 8 + (o >> 3), 0, h1 - 1) * s1 + clampi(bx * 8 + (o & 7), 0, w1 - 1)];
    for (int o = lane; o < 144; o += 32) {
      const int r = o / 12, c = o % 12;
      sm.ref1[warp][r * 14 + c] = ref1[(size_t)clampi(by * 8 + py - 2 + r, 0, h1 - 1) * s1 + clampi(bx * 8 + px - 2 + c, 0, w1 - 1)];
    }
  }
  __syncwarp();
  int dy, dx, qy = 0, qx = 0;
  refine25<8, 14, false>(sm.cur1[warp], sm.ref1[warp], xo1, lane, P.lambda >> 2, 0, &dy, &dx, &qy, &qx);
  const int qy0 = 2 * (py + dy), qx0 = 2 * (px + dx);
  // ---- L0: 16x16 block at (16bx, 16by), +-2 around (qx, qy) ----
  const bool in0 = by * 16 + 16 <= P.height && bx * 16 + 16 <= P.width && by * 16 + qy0 - 2 >= 0 && by * 16 + qy0 + 18 <= P.height &&
                   bx * 16 + qx0 - 2 >= 0 && bx * 16 + qx0 + 18 <= P.width && bx * 16 + qx0 + 20 <= P.stride0;
  int xo0 = 0;
  if (in0) {
    // lane (r, h): 8 samples of row r as one 16-byte load
    *reinterpret_cast<uint4*>(&sm.cur0[warp][(lane >> 1) * 16 + 8 * (lane & 1)]) =
        *reinterpret_cast<const uint4*>(cur0 + (size_t)(by * 16 + (lane >> 1)) * P.stride0 + bx * 16 + 8 * (lane & 1));
    // 20 rows of 20 samples as aligned words (11 per row cover an odd start): lane = (row of two, word)
    const int wx = bx * 16 + qx0 - 2;
    xo0 = wx & 1;
    const int wr = lane >> 4, wc = lane & 15;
    if (wc < 11) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(ref0 + (size_t)(by * 16 + qy0 - 2 + wr) * P.stride0 + (wx - xo0)) + wc;
      uint32_t* dst = reinterpret_cast<uint32_t*>(sm.ref0[warp]) + wr * (kRS0 / 2) + wc;
#pragma unroll
      for (int i = 0; i < 10; i++) dst[i * kRS0] = src[(size_t)i * P.stride0];   // rows wr, wr + 2, ...: two rows = stride0 words
    }
  } else {
    for (int o = lane; o < 256; o += 32)
      sm.cur0[warp][o] = cur0[(size_t)clampi(by * 16 + (o >> 4), 0, P.height - 1) * P.stride0 + clampi(bx * 16 + (o & 15), 0, P.width - 1)];
    for (int o = lane; o < 400; o += 32) {
      const int r = o / 20, c = o % 20;
      sm.ref0[warp][r * kRS0 + c] =
          ref0[(size_t)clampi(by * 16 + qy0 - 2 + r, 0, P.height - 1) * P.stride0 + clampi(bx * 16 + qx0 - 2 + c, 0, P.width - 1)];
    }
  }
  __syncwarp();
  refine25<16, kRS0, true>(sm.cur0[warp], sm.ref0[warp], xo0, lane, P.lambda, P.lambda, &dy, &dx, &qy, &qx);
  if (lane < 4) {
    const int uy = by * 2 + (lane >> 1), ux = bx * 2 + (lane & 1);
    const int w8 = P.width >> 3, h8 = P.height >> 3;
    if (uy < h8 && ux < w8) {
      int16_t* out = P.mv_out + ((size_t)frame * w8 * h8 + (size_t)uy * w8 + ux) * 2;
      out[0] = (int16_t)((qy0 + dy) * 8 + 2 * qy);
      out[1] = (int16_t)((qx0 + dx) * 8 + 2 * qx);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Vector-field regularisation (oracle: orc_mv_dominant, orc_me_sbrd).
//   hme_gather_kernel : 16x16 block vectors out of the per-8x8 vector plane + histogram of the field (1024 hashed
//                       bins: count, largest packed vector) for the dominant vector.
//   hme_sbrd_kernel   : one half sweep of the superblock-level rate-distortion regularisation (below).
__device__ __forceinline__ uint32_t mv_pack(int mvy, int mvx) { return ((uint32_t)(uint16_t)(int16_t)mvy << 16) | (uint16_t)(int16_t)mvx; }
__device__ __forceinline__ uint32_t mv_hash(uint32_t k) { return ((k * 2654435761u) >> 22) & 1023u; }

__global__ void __launch_bounds__(256) hme_gather_kernel(const HmeLaunch P, const int16_t* __restrict__ field8, int16_t* field16,
                                                         int from8) {
  // from8 = 1: read the per-8x8 plane (after the refine kernel); 0: field16 already holds the block vectors
  const int n1x = (P.width + 15) / 16, n1y = (P.height + 15) / 16, n1 = n1x * n1y;
  const int frame = blockIdx.z, b = blockIdx.x * 256 + threadIdx.x;
  if (b >= n1) return;
  int16_t* f16 = field16 + ((size_t)frame * n1 + b) * 2;
  int mvy, mvx;
  if (from8) {
    const int w8 = P.width >> 3, h8 = P.height >> 3, bx = b % n1x, by = b / n1x;
    const int16_t* m = field8 + ((size_t)frame * w8 * h8 + (size_t)(by * 2) * w8 + bx * 2) * 2;
    mvy = m[0]; mvx = m[1];
    f16[0] = (int16_t)mvy; f16[1] = (int16_t)mvx;
  } else {
    mvy = f16[0]; mvx = f16[1];
  }
  const uint32_t k = mv_pack(mvy, mvx), h = mv_hash(k);
  uint32_t* hist = P.hist + (size_t)frame * 2048;
  atomicAdd(&hist[h], 1u);
  atomicMax(&hist[1024 + h], k ^ 0x80008000u);
}

// dominant vector of the frame: the fullest bin (lowest index on ties) -> hist[0..1] of the frame's second half is
// left untouched; the result goes to dom[frame] as a packed vector
__global__ void __launch_bounds__(1024) hme_dominant_kernel(const HmeLaunch P, uint32_t* dom) {
  __shared__ unsigned long long best[32];
  const int frame = blockIdx.x, t = threadIdx.x;
  uint32_t* hist = P.hist + (size_t)frame * 2048;
  // key: count high, then the LOWER bin wins -> (count << 10) | (1023 - bin), maximised
  unsigned long long key = ((unsigned long long)hist[t] << 10) | (unsigned)(1023 - t);
  for (int o = 16; o; o >>= 1) key = max(key, __shfl_xor_sync(0xffffffffu, key, o));
  if ((t & 31) == 0) best[t >> 5] = key;
  __syncthreads();
  if (t < 32) {
    key = best[t];
    for (int o = 16; o; o >>= 1) key = max(key, __shfl_xor_sync(0xffffffffu, key, o));
    if (t == 0) dom[frame] = hist[1024 + (1023 - (int)(key & 1023))] ^ 0x80008000u;
  }
  __syncthreads();
  // clear for the next sweep
  hist[t] = 0; hist[1024 + t] = 0;
}

// Superblock-level rate-distortion regularisation (oracle: orc_me_sbrd).  One CTA of 8 warps per 64x64 superblock of
// the checkerboard colour of this half sweep (8 warps: the decisions are taken by one warp, so small CTAs keep more
// superblocks in flight per SM); warp w owns the 16x16 blocks w and w + 8 (block b = (b >> 2, b & 3)).  For every candidate vector the
// CTA stages the 65x65 window of the reference SOURCE picture the superblock points at (aligned 32-bit words, 68-sample
// row pitch, the odd start handled when the words are read), lane (i, h) of a warp computes the bilinear quarter-sample
// SAD of 8 samples of row i of its block and the warp adds them up: T[candidate][block].  Warp 0 then takes the
// decisions with one lane per candidate: relaxation of the blocks in raster order, the 64x64 merge test, the 32x32 tests.
// Superblocks of one colour share no edge, so the field is updated in place.
constexpr int kSbCand = 22;
constexpr int kWinPitch = 68;   // samples (34 words)
struct SbrdSmem {
  uint32_t win[2][65 * kWinPitch / 2];   // two candidate windows per round
  int T[kSbCand][16];
  uint32_t cand[kSbCand];
  uint32_t loc[6][6];        // vectors of the superblock's blocks and the blocks around it ([r + 1][c + 1]); 0xFFFFFFFF outside
  uint8_t have[6][6];        // that block lies inside the picture
  int li[16];                // candidate index of the vector a block holds
  int ncand;
};

__device__ __forceinline__ uint32_t hw(const uint32_t* w, int i) { return (i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xFFFFu); }

__global__ void __launch_bounds__(256) hme_sbrd_kernel(const HmeLaunch P, uint32_t* field16, const uint32_t* __restrict__ dom,
                                                       int colour, int16_t* field8) {
  __shared__ SbrdSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int frame = blockIdx.z;
  const int W = P.width, H = P.height;
  const int n1x = (W + 15) / 16, n1y = (H + 15) / 16, n1 = n1x * n1y;
  const int nsx = (n1x + 3) / 4;
  const int sby = blockIdx.y, sbx = 2 * blockIdx.x + ((sby + colour) & 1);
  if (sbx >= nsx) return;
  const int y0 = sby * 4, x0 = sbx * 4, nr = min(4, n1y - y0), nc = min(4, n1x - x0);
  uint32_t* F = field16 + (size_t)frame * n1;
  const uint16_t* cur0 = P.cur[0] + (size_t)P.cur_slot[frame] * P.elems0;
  const uint16_t* ref0 = P.ref[0] + (size_t)P.ref_slot[frame] * P.elems0;
  // ---- the field around the superblock, the candidate list ----
  if (tid < 36) {
    const int r = tid / 6 - 1, c = tid % 6 - 1;
    const int by = y0 + r, bx = x0 + c;
    const bool in = by >= 0 && by < n1y && bx >= 0 && bx < n1x && r <= nr && c <= nc;
    sm.loc[r + 1][c + 1] = in ? F[by * n1x + bx] : 0xFFFFFFFFu;
    sm.have[r + 1][c + 1] = in ? 1 : 0;
  }
  __syncthreads();
  if (warp == 0) {
    // entry i of the ordered list: zero, dominant, left, above, right, below, own vectors in raster order
    uint32_t k = 0; bool valid = false;
    if (lane == 0) { k = 0; valid = true; }
    else if (lane == 1) { const uint32_t d = dom[frame]; k = (d >> 16) | (d << 16); valid = true; }   // mv_pack -> (col << 16) | row
    else if (lane == 2) { valid = x0 > 0; k = sm.loc[1][0]; }
    else if (lane == 3) { valid = y0 > 0; k = sm.loc[0][1]; }
    else if (lane == 4) { valid = x0 + nc < n1x; k = sm.loc[1][nc + 1]; }
    else if (lane == 5) { valid = y0 + nr < n1y; k = sm.loc[nr + 1][1]; }
    else if (lane < 22) { const int b = lane - 6, r = b >> 2, c = b & 3; valid = r < nr && c < nc; k = sm.loc[r + 1][c + 1]; }
    bool first = valid;
    for (int j = 0; j < 21; j++) {
      const uint32_t kj = __shfl_sync(0xffffffffu, k, j);
      const bool vj = __shfl_sync(0xffffffffu, (int)valid, j) != 0;
      if (j < lane && vj && kj == k) first = false;
    }
    const unsigned m = __ballot_sync(0xffffffffu, first);
    if (first) sm.cand[__popc(m & ((1u << lane) - 1))] = k;
    if (lane == 0) sm.ncand = __popc(m);
  }
  // ---- this lane's 8 samples of a row of each of the warp's two blocks (warp w: blocks w and w + 8; clamped at the picture edge) ----
  const int li = lane >> 1, lh = lane & 1;
  unsigned c8[2][8];
  bool blk_on[2];
#pragma unroll
  for (int hb = 0; hb < 2; hb++) {
    const int blk = warp + 8 * hb, br = blk >> 2, bc = blk & 3;
    blk_on[hb] = br < nr && bc < nc;
    const int y = (y0 + br) * 16 + li, x = (x0 + bc) * 16 + 8 * lh;
    if (blk_on[hb] && y < H && x + 8 <= W) {
      const uint4 v = *reinterpret_cast<const uint4*>(cur0 + (size_t)y * P.stride0 + x);
      c8[hb][0] = v.x & 0xFFFFu; c8[hb][1] = v.x >> 16; c8[hb][2] = v.y & 0xFFFFu; c8[hb][3] = v.y >> 16;
      c8[hb][4] = v.z & 0xFFFFu; c8[hb][5] = v.z >> 16; c8[hb][6] = v.w & 0xFFFFu; c8[hb][7] = v.w >> 16;
    } else {
#pragma unroll
      for (int j = 0; j < 8; j++) c8[hb][j] = blk_on[hb] ? cur0[(size_t)clampi(y, 0, H - 1) * P.stride0 + clampi(x + j, 0, W - 1)] : 0;
    }
  }
  __syncthreads();
  const int ncand = sm.ncand;
  if (ncand == 1) {
    // every block of the superblock and everything around it holds the zero vector: nothing can change
    if (field8 && tid < 16) {
      const int r = tid >> 2, c = tid & 3;
      if (r < nr && c < nc) {
        const int w8 = W >> 3, h8 = H >> 3;
        for (int u = 0; u < 4; u++) {
          const int uy = (y0 + r) * 2 + (u >> 1), ux = (x0 + c) * 2 + (u & 1);
          if (uy < h8 && ux < w8) reinterpret_cast<uint32_t*>(field8)[(size_t)frame * w8 * h8 + (size_t)uy * w8 + ux] = 0u;
        }
      }
    }
    return;
  }
  // two candidates per round: both windows are loaded back to back (their global-load latencies overlap), one barrier
  // pair serves two SAD passes
  auto stage_window = [&](int k, uint32_t* win) {
    const uint32_t cv = sm.cand[k];
    const int mvy = (int16_t)(cv & 0xFFFFu), mvx = (int16_t)(cv >> 16);
    const int wy = y0 * 16 + (mvy >> 3), wx = x0 * 16 + (mvx >> 3);   // window origin in the picture
    const int wxa = wx & ~1;                                           // first staged column (even)
    if (wy >= 0 && wy + 65 <= H && wxa >= 0 && wx + 65 <= W && wxa + kWinPitch <= P.stride0) {
      // warp w stages rows w, w + 8, ..., w + 56 (and warp 0 row 64): lane = word, lanes 0 and 1 also words 32 and 33
      const uint32_t* src = reinterpret_cast<const uint32_t*>(ref0 + (size_t)(wy + warp) * P.stride0 + wxa) + lane;
      uint32_t* dst = win + warp * (kWinPitch / 2) + lane;
      const size_t step = (size_t)4 * P.stride0;   // 8 rows in words
#pragma unroll
      for (int i = 0; i < 8; i++) {
        dst[i * 8 * (kWinPitch / 2)] = src[i * step];
        if (lane < 2) dst[i * 8 * (kWinPitch / 2) + 32] = src[i * step + 32];
      }
      if (warp == 0) {
        dst[64 * (kWinPitch / 2)] = src[8 * step];
        if (lane < 2) dst[64 * (kWinPitch / 2) + 32] = src[8 * step + 32];
      }
    } else {
      for (int o = tid; o < 65 * (kWinPitch / 2); o += 256) {
        const int r = o / (kWinPitch / 2), c = o - r * (kWinPitch / 2);
        const uint16_t* row = ref0 + (size_t)clampi(wy + r, 0, H - 1) * P.stride0;
        win[o] = (uint32_t)row[clampi(wxa + 2 * c, 0, W - 1)] | ((uint32_t)row[clampi(wxa + 2 * c + 1, 0, W - 1)] << 16);
      }
    }
  };
  auto sad_pass = [&](int k, const uint32_t* win) {
    const uint32_t cv = sm.cand[k];
    const int mvy = (int16_t)(cv & 0xFFFFu), mvx = (int16_t)(cv >> 16);
    const int fx = (mvx & 7) >> 1, fy = (mvy & 7) >> 1;
    const int xo = (x0 * 16 + (mvx >> 3)) & 1;              // offset of the window in its staged rows
    const int w00 = (4 - fx) * (4 - fy), w01 = fx * (4 - fy), w10 = (4 - fx) * fy, w11 = fx * fy;
#pragma unroll
    for (int hb = 0; hb < 2; hb++) {
      if (!blk_on[hb]) continue;   // warp-uniform
      const int blk = warp + 8 * hb, br = blk >> 2, bc = blk & 3;
      const uint32_t* r0 = &win[(br * 16 + li) * (kWinPitch / 2) + bc * 8 + 4 * lh];
      uint32_t a[5], b[5];
#pragma unroll
      for (int j = 0; j < 5; j++) { a[j] = r0[j]; b[j] = r0[kWinPitch / 2 + j]; }
      unsigned sad = 0;
      if ((fx | fy) == 0) {   // whole-sample vector: the interpolation is the identity
        if (xo == 0) {
#pragma unroll
          for (int j = 0; j < 8; j++) sad = __usad(c8[hb][j], hw(a, j), sad);
        } else {
#pragma unroll
          for (int j = 0; j < 8; j++) sad = __usad(c8[hb][j], hw(a, j + 1), sad);
        }
      } else if (xo == 0) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const unsigned p = (unsigned)(w00 * (int)hw(a, j) + w01 * (int)hw(a, j + 1) + w10 * (int)hw(b, j) + w11 * (int)hw(b, j + 1) + 8) >> 4;
          sad = __usad(c8[hb][j], p, sad);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const unsigned p = (unsigned)(w00 * (int)hw(a, j + 1) + w01 * (int)hw(a, j + 2) + w10 * (int)hw(b, j + 1) + w11 * (int)hw(b, j + 2) + 8) >> 4;
          sad = __usad(c8[hb][j], p, sad);
        }
      }
      sad = __reduce_add_sync(0xffffffffu, sad);
      if (lane == 0) sm.T[k][blk] = (int)sad;
    }
  };
  for (int k = 0; k < ncand; k += 2) {
    __syncthreads();                                       // the windows of the round before have been read
    stage_window(k, sm.win[0]);
    if (k + 1 < ncand) stage_window(k + 1, sm.win[1]);
    __syncthreads();
    sad_pass(k, sm.win[0]);
    if (k + 1 < ncand) sad_pass(k + 1, sm.win[1]);
  }
  __syncthreads();
  if (warp != 0) return;
  // ---- decisions: lane = candidate ----
  const uint32_t myc = lane < ncand ? sm.cand[lane] : 0;
  const unsigned kInf = 0xFFFFFFFFu;
  for (int b = 0; b < 16; b++) {
    const int r = b >> 2, c = b & 3;
    if (r >= nr || c >= nc) continue;
    unsigned key = kInf;
    if (lane < ncand) {
      const int diff = (sm.have[r + 1][c] && sm.loc[r + 1][c] != myc) + (sm.have[r + 1][c + 2] && sm.loc[r + 1][c + 2] != myc) +
                       (sm.have[r][c + 1] && sm.loc[r][c + 1] != myc) + (sm.have[r + 2][c + 1] && sm.loc[r + 2][c + 1] != myc);
      key = ((unsigned)(sm.T[lane][b] + P.lam_s * diff) << 5) | (unsigned)lane;
    }
    key = __reduce_min_sync(0xffffffffu, key);
    const int bk = key & 31;
    __syncwarp();
    if (lane == 0) { sm.loc[r + 1][c + 1] = sm.cand[bk]; sm.li[b] = bk; }
    __syncwarp();
  }
  auto test = [&](int r0, int c0, int r1, int c1) -> bool {
    // the field as it stands: lane = block
    int jc = 0;
    {
      const int r = lane >> 2, c = lane & 3;
      if (lane < 16 && r >= r0 && r < r1 && c >= c0 && c < c1) {
        const uint32_t k = sm.loc[r + 1][c + 1];
        const bool same = (sm.have[r + 1][c] && sm.loc[r + 1][c] == k) || (sm.have[r][c + 1] && sm.loc[r][c + 1] == k);
        jc = sm.T[sm.li[lane]][lane] + P.lam_r * (same ? 1 : 12);
      }
    }
    jc = (int)__reduce_add_sync(0xffffffffu, (unsigned)jc);
    unsigned key = kInf;
    if (lane < ncand) {
      const bool nb = myc == 0 || (sm.have[r0 + 1][c0] && sm.loc[r0 + 1][c0] == myc) || (sm.have[r0][c0 + 1] && sm.loc[r0][c0 + 1] == myc);
      int j = P.lam_r * (nb ? 1 : 12);
      for (int r = r0; r < r1; r++) for (int c = c0; c < c1; c++) j += sm.T[lane][r * 4 + c];
      key = ((unsigned)j << 5) | (unsigned)lane;
    }
    key = __reduce_min_sync(0xffffffffu, key);
    const bool take = (int)(key >> 5) < jc;
    __syncwarp();
    if (take) {
      const int r = lane >> 2, c = lane & 3;
      if (lane < 16 && r >= r0 && r < r1 && c >= c0 && c < c1) { sm.loc[r + 1][c + 1] = sm.cand[key & 31]; sm.li[lane] = key & 31; }
    }
    __syncwarp();
    return take;
  };
  if (!test(0, 0, nr, nc))
    for (int qr = 0; qr < nr; qr += 2)
      for (int qc = 0; qc < nc; qc += 2) test(qr, qc, min(qr + 2, nr), min(qc + 2, nc));
  if (lane < 16) {
    const int r = lane >> 2, c = lane & 3;
    if (r < nr && c < nc) {
      const uint32_t k = sm.loc[r + 1][c + 1];
      F[(y0 + r) * n1x + x0 + c] = k;
      if (field8) {
        const int w8 = W >> 3, h8 = H >> 3;
        for (int u = 0; u < 4; u++) {
          const int uy = (y0 + r) * 2 + (u >> 1), ux = (x0 + c) * 2 + (u & 1);
          if (uy < h8 && ux < w8) reinterpret_cast<uint32_t*>(field8)[(size_t)frame * w8 * h8 + (size_t)uy * w8 + ux] = k;
        }
      }
    }
  }
}

}  // namespace

cudaError_t launch_hme_sbrd(const HmeLaunch& p, int n, cudaStream_t s) {
  if (p.lam_s <= 0 || p.sbrd_passes <= 0) return cudaSuccess;
  const int n1x = (p.width + 15) / 16, n1y = (p.height + 15) / 16, n1 = n1x * n1y;
  const int nsx = (n1x + 3) / 4, nsy = (n1y + 3) / 4;
  uint32_t* field = reinterpret_cast<uint32_t*>(p.mv_tmp);
  uint32_t* dom = p.hist + (size_t)n * 2048;
  cudaError_t e = cudaMemsetAsync(p.hist, 0, (size_t)n * 2048 * sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  hme_gather_kernel<<<dim3((n1 + 255) / 256, 1, n), 256, 0, s>>>(p, p.mv_out, p.mv_tmp, 1);
  hme_dominant_kernel<<<n, 1024, 0, s>>>(p, dom);
  const dim3 grid((nsx + 1) / 2, nsy, n);
  for (int half = 0; half < 2 * p.sbrd_passes; half++)
    hme_sbrd_kernel<<<grid, 256, 0, s>>>(p, field, dom, half & 1, half + 1 == 2 * p.sbrd_passes || half + 2 == 2 * p.sbrd_passes ? p.mv_out : nullptr);
  return cudaGetLastError();
}

cudaError_t launch_pyramid(const uint16_t* l0, uint16_t* l1, uint16_t* l2, int stride0, int rows0, size_t elems0,
                           int n_frames, cudaStream_t s) {
  const size_t total = (size_t)(stride0 >> 2) * (rows0 >> 2) * n_frames;
  pyramid_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(l0, l1, l2, stride0, rows0, elems0, n_frames);
  return cudaGetLastError();
}

cudaError_t launch_hme(const HmeLaunch& p, int n_frames, cudaStream_t s) {
  const int n2x = (p.width + 31) / 32, n2y = (p.height + 31) / 32;
  dim3 g2((n2x + 3) / 4, (n2y + 3) / 4, n_frames);
  hme_l2_kernel<<<g2, 256, 0, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int n1 = ((p.width + 15) / 16) * ((p.height + 15) / 16);
  dim3 g1((n1 + 7) / 8, 1, n_frames);
  hme_refine_kernel<<<g1, 256, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace av1b
