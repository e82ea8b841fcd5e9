// Motion-compensated temporal filter of key / anchor SOURCE pictures for sm_100a (oracle: orc_mctf).
//
// One warp per 16x16 luma block of the picture that is filtered.  For every neighbour in time the warp stages the
// (16+7)^2 luma window (and the two (8+7)^2 chroma windows) the block's vector points at, runs the normative separable
// 8-tap interpolation (spec 7.11.3.4: the same arithmetic the inter prediction uses), derives the block weight from the
// luma mean squared error and the sample weights from the sample differences, and accumulates weighted sums in
// registers; the picture and every neighbour are read once, the filtered picture is written once:
//   algorithmic bytes = (2 + n_nb) * S  per filtered picture.
//
// Replaces arithmetic the reference delegates to av1an + SVT-AV1 (/root/reference/crates/daemon/src/encode/av1an.rs:14
// `--film-grain 20` = denoise + synthesis, `--lookahead 40`).
#include <cuda_runtime.h>
#include <stdint.h>
#include "av1_tables_dev.cuh"
#include "kernels.cuh"

namespace av1b {
namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

constexpr int kWarps = 8;
constexpr int kWinPitch = 26;        // samples per staged row (13 words): 23 + an odd start
struct MctfSmem {
  uint32_t win[kWarps][23 * kWinPitch / 2];   // staged reference window (luma 23x23, chroma 15x15), rows as aligned words
  int32_t mid[kWarps][23 * 17];               // after the horizontal pass (pitch N + 1)
  uint16_t pred[kWarps][16 * 16];
};

// Normative prediction of an N x N block at (x, y) of a plane from `ref` with the vector (mv_row, mv_col) in 1/8 luma
// samples (oracle: orc_inter_predict): all lanes of the warp take part, result in pred[N*N].
//   whole-sample vectors copy the block (both filter passes are the identity then);
//   windows inside the picture are staged as aligned 32-bit words without clamping (an odd start is an offset at read time);
//   a lane filters four neighbouring outputs from eleven inputs held in registers.
template <int N>
__device__ __forceinline__ void mc_block(const uint16_t* __restrict__ ref, int stride, int pw, int ph, int x, int y, int mv_row,
                                         int mv_col, int ss, int bd, int lane, uint32_t* winw, int32_t* mid, uint16_t* pred) {
  const int x16 = (x << 4) + ((2 * mv_col) >> ss), y16 = (y << 4) + ((2 * mv_row) >> ss);
  const int ix = x16 >> 4, iy = y16 >> 4, fx = x16 & 15, fy = y16 & 15;
  constexpr int WN = N + 7, MP = N + 1;
  if ((fx | fy) == 0) {
    if (ix >= 0 && iy >= 0 && ix + N <= pw && iy + N <= ph) {
      for (int o = lane; o < N * N; o += 32) pred[o] = ref[(size_t)(iy + o / N) * stride + ix + o % N];
    } else {
      for (int o = lane; o < N * N; o += 32) pred[o] = ref[(size_t)clampi(iy + o / N, 0, ph - 1) * stride + clampi(ix + o % N, 0, pw - 1)];
    }
    __syncwarp();
    return;
  }
  uint16_t* win = reinterpret_cast<uint16_t*>(winw);
  const int wx = ix - 3, wy = iy - 3;
  int xo = 0;
  if (wx >= 1 && wy >= 0 && wx + WN + 1 <= pw && wy + WN <= ph) {
    xo = wx & 1;
    constexpr int WW = (WN + 2) / 2;   // words per row that cover WN samples from an even or odd start
    const uint32_t* base = reinterpret_cast<const uint32_t*>(ref + (size_t)wy * stride + (wx - xo));
    for (int o = lane; o < WN * WW; o += 32) {
      const int r = o / WW, c = o - r * WW;
      winw[r * (kWinPitch / 2) + c] = base[(size_t)r * (stride >> 1) + c];
    }
  } else {
    for (int o = lane; o < WN * WN; o += 32) {
      const int r = o / WN, c = o - r * WN;
      win[r * kWinPitch + c] = ref[(size_t)clampi(wy + r, 0, ph - 1) * stride + clampi(wx + c, 0, pw - 1)];
    }
  }
  __syncwarp();
  // horizontal pass: task = (row, group of four outputs)
  {
    int kx[8];
#pragma unroll
    for (int t = 0; t < 8; t++) kx[t] = tbl::sub_pel_filters_8[fx][t];
    constexpr int G = N / 4;
    for (int o = lane; o < WN * G; o += 32) {
      const int r = o / G, c0 = (o - r * G) * 4;
      const uint16_t* wp = win + r * kWinPitch + xo + c0;
      int v[11];
#pragma unroll
      for (int j = 0; j < 11; j++) v[j] = wp[j];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        int sacc = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) sacc += kx[t] * v[j + t];
        mid[r * MP + c0 + j] = (sacc + 4) >> 3;
      }
    }
  }
  __syncwarp();
  // vertical pass: task = (column, group of four output rows)
  {
    int ky[8];
#pragma unroll
    for (int t = 0; t < 8; t++) ky[t] = tbl::sub_pel_filters_8[fy][t];
    const int maxv = (1 << bd) - 1;
    constexpr int G = N / 4;
    for (int o = lane; o < N * G; o += 32) {
      const int c = o % N, r0 = (o / N) * 4;
      int v[11];
#pragma unroll
      for (int j = 0; j < 11; j++) v[j] = mid[(r0 + j) * MP + c];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        int sacc = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) sacc += ky[t] * v[j + t];
        pred[(r0 + j) * N + c] = (uint16_t)clampi((sacc + 1024) >> 11, 0, maxv);
      }
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kWarps * 32, 3) mctf_kernel(const MctfLaunch P) {
  __shared__ MctfSmem sm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const Av1bGeom& g = P.g;
  const int W = g.width, H = g.height;
  const int n1x = (W + 15) / 16, n1y = (H + 15) / 16;
  const int blk = blockIdx.x * kWarps + warp;
  if (blk >= n1x * n1y) return;
  const int bx = blk % n1x, by = blk / n1x;
  const int bw = min(16, W - bx * 16), bh = min(16, H - by * 16);
  uint32_t* win = sm.win[warp];
  int32_t* mid = sm.mid[warp];
  uint16_t* pred = sm.pred[warp];
  // lane (r, h): luma row r, columns 8h .. 8h+7; chroma: samples 2 lane, 2 lane + 1 of the 8x8 block
  const int lr = lane >> 1, lc = (lane & 1) * 8;
  uint32_t numy[8], deny[8], numc[2][2], denc[2][2];
  int cury[8], curc[2][2];
  {
    const uint16_t* cp = P.cur[0] + (size_t)(by * 16 + lr) * g.stride[0] + bx * 16 + lc;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const bool in = lr < bh && lc + j < bw;
      cury[j] = in ? cp[j] : 0;
      numy[j] = 256u * (uint32_t)cury[j]; deny[j] = 256u;
    }
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int o = 2 * lane + j, r = o >> 3, c = o & 7;
        const bool in = r < (bh >> 1) && c < (bw >> 1);
        curc[p][j] = in ? P.cur[1 + p][(size_t)(by * 8 + r) * g.stride[1] + bx * 8 + c] : 0;
        numc[p][j] = 256u * (uint32_t)curc[p][j]; denc[p][j] = 256u;
      }
  }
  for (int k = 0; k < P.n_nb; k++) {
    const int16_t* mv = P.mvs[k] + ((size_t)(by * 2) * g.w8 + bx * 2) * 2;
    const int mvr = mv[0], mvc = mv[1];
    mc_block<16>(P.nb[k][0], g.stride[0], W, H, bx * 16, by * 16, mvr, mvc, 0, P.bit_depth, lane, win, mid, pred);
    int py[8];
    unsigned sse = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      py[j] = pred[lr * 16 + lc + j];
      const int d = (lr < bh && lc + j < bw) ? py[j] - cury[j] : 0;
      sse += (unsigned)(d * d);
    }
    for (int o = 16; o; o >>= 1) sse += __shfl_xor_sync(0xffffffffu, sse, o);
    const int mse = (int)(sse / (unsigned)(bw * bh));
    const int wb = clampi(16 - (int)(((long long)16 * mse) / P.thr_b), 0, 16);
    if (wb == 0) continue;   // warp-uniform
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int d = py[j] - cury[j];
      const int w = wb * clampi(16 - (16 * d * d) / P.thr_p, 0, 16);
      numy[j] += (uint32_t)w * (uint32_t)py[j]; deny[j] += (uint32_t)w;
    }
#pragma unroll
    for (int p = 0; p < 2; p++) {
      __syncwarp();
      mc_block<8>(P.nb[k][1 + p], g.stride[1], W >> 1, H >> 1, bx * 8, by * 8, mvr, mvc, 1, P.bit_depth, lane, win, mid, pred);
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int pv = pred[2 * lane + j];
        const int d = pv - curc[p][j];
        const int w = wb * clampi(16 - (16 * d * d) / P.thr_p, 0, 16);
        numc[p][j] += (uint32_t)w * (uint32_t)pv; denc[p][j] += (uint32_t)w;
      }
    }
    __syncwarp();
  }
  {
    uint16_t* op = P.out[0] + (size_t)(by * 16 + lr) * g.stride[0] + bx * 16 + lc;
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (lr < bh && lc + j < bw) op[j] = (uint16_t)((numy[j] + deny[j] / 2) / deny[j]);
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int o = 2 * lane + j, r = o >> 3, c = o & 7;
        if (r < (bh >> 1) && c < (bw >> 1))
          P.out[1 + p][(size_t)(by * 8 + r) * g.stride[1] + bx * 8 + c] = (uint16_t)((numc[p][j] + denc[p][j] / 2) / denc[p][j]);
      }
  }
}

// Noise level of a source picture (oracle: orc_noise_estimate): one CTA per 64x64 superblock = 16 blocks of 16x16, 16
// lanes per block (lane j = row j of the block); |I * N| over the 14x14 inner samples, block sums into a 4096-bin histogram.
__global__ void __launch_bounds__(256) noise_hist_kernel(Av1bGeom g, const uint16_t* __restrict__ src_y, uint32_t* hist) {
  const int tid = threadIdx.x, blk = tid >> 4, j = tid & 15;
  const int bx = blockIdx.x * 4 + (blk & 3), by = blockIdx.y * 4 + (blk >> 2);
  const bool inside = bx * 16 + 16 <= g.width && by * 16 + 16 <= g.height;
  int sum = 0;
  if (inside && j >= 1 && j <= 14) {
    const uint16_t* p = src_y + (size_t)(by * 16 + j) * g.stride[0] + bx * 16;
    const int st = g.stride[0];
    int a0 = p[-st], a1 = p[-st + 1], b0 = p[0], b1 = p[1], c0 = p[st], c1 = p[st + 1];
#pragma unroll
    for (int x = 1; x < 15; x++) {
      const int a2 = p[-st + x + 1], b2 = p[x + 1], c2 = p[st + x + 1];
      sum += abs(a0 - 2 * a1 + a2 - 2 * b0 + 4 * b1 - 2 * b2 + c0 - 2 * c1 + c2);
      a0 = a1; a1 = a2; b0 = b1; b1 = b2; c0 = c1; c1 = c2;
    }
  }
#pragma unroll
  for (int o = 8; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (inside && j == 0) atomicAdd(&hist[min(sum >> 4, 4095)], 1u);
}

// Scene-change score of every picture of a batch against the picture before it (oracle: orc_scene_score): sum of absolute
// luma differences on the 1/8 x 1/8 grid that starts at sample (4, 4) -- the thumbnails the drop-in executable cuts its chunks
// by, computed where the pictures already are.  Picture 0 is compared with `prev` (the last picture of the batch before; no
// score when there is none).
__global__ void __launch_bounds__(256) scene_score_kernel(Av1bGeom g, const uint16_t* __restrict__ src_y, size_t elems0,
                                                          const uint16_t* __restrict__ prev, uint32_t* score) {
  const int b = blockIdx.y;
  const int tw = (g.width - 4 + 7) / 8, th = (g.height - 4 + 7) / 8;
  const int i = blockIdx.x * 256 + threadIdx.x;
  const uint16_t* cur = src_y + (size_t)b * elems0;
  const uint16_t* old = b ? cur - elems0 : prev;
  unsigned d = 0;
  if (i < tw * th && old) {
    const size_t o = (size_t)(4 + 8 * (i / tw)) * g.stride[0] + 4 + 8 * (i % tw);
    d = (unsigned)abs((int)cur[o] - (int)old[o]);
  }
  d = __reduce_add_sync(0xffffffffu, d);
  if ((threadIdx.x & 31) == 0 && d) atomicAdd(&score[b], d);
}

}  // namespace

cudaError_t launch_scene_score(const Av1bGeom& g, const uint16_t* src_y, size_t elems0, const uint16_t* prev, int n_frames,
                               uint32_t* score, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(score, 0, (size_t)n_frames * sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  const int tw = (g.width - 4 + 7) / 8, th = (g.height - 4 + 7) / 8;
  scene_score_kernel<<<dim3((tw * th + 255) / 256, n_frames), 256, 0, s>>>(g, src_y, elems0, prev, score);
  return cudaGetLastError();
}

cudaError_t launch_noise_hist(const Av1bGeom& g, const uint16_t* src_y, uint32_t* hist, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(hist, 0, 4096 * sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  dim3 grid(g.sb_cols, g.sb_rows);
  noise_hist_kernel<<<grid, 256, 0, s>>>(g, src_y, hist);
  return cudaGetLastError();
}

cudaError_t launch_mctf(const MctfLaunch& p, cudaStream_t s) {
  const int n1 = ((p.g.width + 15) / 16) * ((p.g.height + 15) / 16);
  mctf_kernel<<<(n1 + kWarps - 1) / kWarps, kWarps * 32, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace av1b
