// Batched normative inverse transform + reconstruction for all 19 AV1 transform sizes (4x4 .. 64x64,
// 2:1 and 4:1 rectangles) and every legal transform type (spec 7.13.3): the "kernel bit-exact suite"
// of BASELINE.json config 2.  One CTA per transform block; the row pass runs one thread per coded
// row (at most 32: AV1 zeroes everything beyond 32x32), the column pass one thread per column.  The
// butterflies are the shared normative graphs of av1_inv_txfm1d.h (integer pipes only).
//
// Replaces arithmetic the reference delegates to av1an + SVT-AV1
// (/root/reference/crates/daemon/src/encode/av1an.rs:126-139).
// Bit-exact against libaom's av1_inv_txfm2d_add_*_c (tests/test_gpu_kernel_suite.py).
#include <cuda_runtime.h>
#include <stdint.h>
#include "av1_inv_txfm1d.h"
#include "av1_tables_dev.cuh"
#include "kernels.cuh"

namespace av1b {
using namespace av1tx;
namespace {

enum { T_DCT = 0, T_ADST = 1, T_FLIP = 2, T_IDT = 3 };
__constant__ uint8_t k_vtype[16] = {T_DCT, T_ADST, T_DCT, T_ADST, T_FLIP, T_DCT, T_FLIP, T_ADST, T_FLIP,
                                    T_IDT, T_DCT, T_IDT, T_ADST, T_IDT, T_FLIP, T_IDT};
__constant__ uint8_t k_htype[16] = {T_DCT, T_DCT, T_ADST, T_ADST, T_DCT, T_FLIP, T_FLIP, T_FLIP, T_ADST,
                                    T_IDT, T_IDT, T_DCT, T_IDT, T_ADST, T_IDT, T_FLIP};

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

__host__ __device__ constexpr int ilog2c(int n) { return n <= 1 ? 0 : 1 + ilog2c(n >> 1); }
// Transform_Row_Shift (spec 7.13.3)
__host__ __device__ constexpr int row_shift_of(int w, int h) {
  const int lw = ilog2c(w), lh = ilog2c(h), s = lw + lh;
  // 4x4:0 4x8/8x4:0 8x8:1 4x16/16x4:1 8x16/16x8:1 16x16:2 8x32/32x8:2 16x32/32x16:1 32x32:2 16x64/64x16:2
  // 32x64/64x32:1 64x64:2
  return s == 4 ? 0 : s == 5 ? 0 : s == 6 ? 1 : s == 7 ? 1 : s == 8 ? 2 : s == 9 ? 1 : s == 10 ? 2 : s == 11 ? 1 : 2;
}

template <int W, int H>
__global__ void __launch_bounds__((W > H ? W : H) < 32 ? 32 : (W > H ? W : H))
inv_txfm_add_kernel(const int32_t* __restrict__ coef, uint16_t* __restrict__ dst, int tx_type, int bd) {
  constexpr int CW = W < 32 ? W : 32, CH = H < 32 ? H : 32;
  constexpr int LW = ilog2c(W), LH = ilog2c(H);
  constexpr bool kRect = (LW - LH == 1) || (LH - LW == 1);
  constexpr int kRowShift = row_shift_of(W, H);
  __shared__ int32_t buf[CH * (W + 1)];
  const int t = threadIdx.x;
  const int32_t* cf = coef + (size_t)blockIdx.x * 1024;
  uint16_t* d = dst + (size_t)blockIdx.x * W * H;
  const int vt = k_vtype[tx_type], ht = k_htype[tx_type];
  const int row_range = bd + 8, col_range = bd + 6 > 16 ? bd + 6 : 16;
  if (t < CH) {
    int32_t x[W];
#pragma unroll
    for (int j = 0; j < W; j++) {
      int32_t v = j < CW ? cf[t * CW + j] : 0;
      if (kRect) v = (int32_t)(((int64_t)v * 2896 + 2048) >> 12);
      x[j] = sat(v, row_range);
    }
    inv_1d<W>(ht, x, row_range);
#pragma unroll
    for (int j = 0; j < W; j++) {
      int32_t v = x[ht == T_FLIP ? W - 1 - j : j];
      if (kRowShift > 0) v = (v + (1 << (kRowShift > 0 ? kRowShift - 1 : 0))) >> kRowShift;
      buf[t * (W + 1) + j] = v;
    }
  }
  __syncthreads();
  if (t < W) {
    int32_t x[H];
#pragma unroll
    for (int i = 0; i < H; i++) x[i] = i < CH ? sat(buf[i * (W + 1) + t], col_range) : 0;
    inv_1d<H>(vt, x, col_range);
    const int maxv = (1 << bd) - 1;
#pragma unroll
    for (int i = 0; i < H; i++) {
      const int32_t v = (x[vt == T_FLIP ? H - 1 - i : i] + 8) >> 4;
      const int o = (int)d[i * W + t] + v;
      d[i * W + t] = (uint16_t)(o < 0 ? 0 : (o > maxv ? maxv : o));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Encoder-side forward transform of every size and type (oracle: orc_fwd_txfm2d): DCT / ADST / flipADST / identity,
// 4x4 .. 64x64 with the 2:1 and 4:1 rectangles, in the exact integer matrix form the encode kernels use --
//   x' = residual << 2 (flipped for flipADST), t = (Fv x' + 2^11) >> 12 down the columns,
//   coefficient = (Fh t * mul + rnd) >> (24 + log2(w h) - Transform_Row_Shift - 4) along the rows (64-bit),
// mul = 4096, or 5793 for 2:1 rectangles; only min(w,32) x min(h,32) coefficients exist (AV1 zeroes the rest).
// One CTA per block: the residual and the intermediate live in shared memory, a thread per output.
__device__ __forceinline__ int fwd_coef(int t, int n, int k, int i) {
  if (t == T_IDT) return k != i ? 0 : (n == 4 ? 5793 : n == 8 ? 8192 : n == 16 ? 11585 : 16384);
  if (t == T_DCT) {
    switch (n) {
      case 4: return tbl::fwd_dct4[k][i];
      case 8: return tbl::fwd_dct8[k][i];
      case 16: return tbl::fwd_dct16[k][i];
      case 32: return tbl::fwd_dct32[k][i];
      default: return tbl::fwd_dct64[k][i];
    }
  }
  return n == 4 ? tbl::fwd_adst4[k][i] : (n == 8 ? tbl::fwd_adst8[k][i] : tbl::fwd_adst16[k][i]);
}

__global__ void __launch_bounds__(256) fwd_txfm_kernel(const int16_t* __restrict__ resid, int32_t* __restrict__ coef, int w, int h,
                                                       int tx_type) {
  __shared__ int16_t r[64 * 64];
  __shared__ int32_t t[32 * 64];
  const int tid = threadIdx.x;
  const int cw = min(w, 32), ch = min(h, 32);
  const int vt = k_vtype[tx_type], ht = k_htype[tx_type];
  const int lw = 31 - __clz(w), lh = 31 - __clz(h);
  const bool rect = abs(lw - lh) == 1;
  const int s = lw + lh;
  const int row_shift = s == 4 ? 0 : s == 5 ? 0 : s == 6 ? 1 : s == 7 ? 1 : s == 8 ? 2 : s == 9 ? 1 : s == 10 ? 2 : s == 11 ? 1 : 2;
  const int sh = 24 + lw + lh - row_shift - 4;
  const long long mul = rect ? 5793 : 4096;
  const int16_t* src = resid + (size_t)blockIdx.x * w * h;
  for (int o = tid; o < w * h; o += 256) r[o] = src[o];
  __syncthreads();
  for (int o = tid; o < ch * w; o += 256) {
    const int k = o / w, j = o - k * w;
    const int jj = ht == T_FLIP ? w - 1 - j : j;
    int32_t acc = 0;
    for (int i = 0; i < h; i++) acc += fwd_coef(vt, h, k, i) * ((int32_t)r[(vt == T_FLIP ? h - 1 - i : i) * w + jj] * 4);
    t[o] = (acc + 2048) >> 12;
  }
  __syncthreads();
  int32_t* dst = coef + (size_t)blockIdx.x * cw * ch;
  for (int o = tid; o < ch * cw; o += 256) {
    const int k = o / cw, l = o - k * cw;
    long long acc = 0;
    for (int j = 0; j < w; j++) acc += (long long)fwd_coef(ht, w, l, j) * t[k * w + j];
    dst[o] = (int32_t)((acc * mul + (1ll << (sh - 1))) >> sh);
  }
}

template <int W, int H>
cudaError_t launch_one(const int32_t* coef, uint16_t* dst, int n, int tx_type, int bd, cudaStream_t s) {
  constexpr int T = (W > H ? W : H) < 32 ? 32 : (W > H ? W : H);
  inv_txfm_add_kernel<W, H><<<n, T, 0, s>>>(coef, dst, tx_type, bd);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_fwd_txfm(const int16_t* resid, int32_t* coef, int n_blocks, int w, int h, int tx_type, cudaStream_t s) {
  fwd_txfm_kernel<<<n_blocks, 256, 0, s>>>(resid, coef, w, h, tx_type);
  return cudaGetLastError();
}

cudaError_t launch_inv_txfm_add(const int32_t* coef, uint16_t* dst, int n_blocks, int w, int h, int tx_type,
                                int bit_depth, cudaStream_t s) {
#define CASE(W, H) if (w == W && h == H) return launch_one<W, H>(coef, dst, n_blocks, tx_type, bit_depth, s)
  CASE(4, 4); CASE(8, 8); CASE(16, 16); CASE(32, 32); CASE(64, 64);
  CASE(4, 8); CASE(8, 4); CASE(8, 16); CASE(16, 8); CASE(16, 32); CASE(32, 16); CASE(32, 64); CASE(64, 32);
  CASE(4, 16); CASE(16, 4); CASE(8, 32); CASE(32, 8); CASE(16, 64); CASE(64, 16);
#undef CASE
  return cudaErrorInvalidValue;
}

}  // namespace av1b
