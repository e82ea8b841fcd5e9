// PSNR / SSIM of the reconstruction against the source, luma plane (progress events of the daemon's job metrics:
// /root/reference/crates/daemon/src/metrics.rs:12-30 `psnr`, `ssim`, which the reference leaves at None / 0 --
// job_executor.rs:117-137).  Reporting only: nothing in the bitstream depends on it.
// One thread per row of an 8x8 block (16-byte loads), eight lanes per block: sums of a, b, a^2, b^2, ab; the squared
// error is exact (integers), SSIM per non-overlapping 8x8 window in float.  Reads 2 Y per frame.
#include <cuda_runtime.h>
#include <stdint.h>
#include "kernels.cuh"

namespace av1b {
namespace {

__global__ void __launch_bounds__(256) quality_kernel(Av1bGeom g, int bit_depth, const uint16_t* __restrict__ rec, const uint16_t* __restrict__ src,
                                                      size_t plane_elems, QualityAcc* out) {
  const int frame = blockIdx.z;
  const uint16_t* a0 = rec + (size_t)frame * plane_elems;
  const uint16_t* b0 = src + (size_t)frame * plane_elems;
  const int t = threadIdx.x, row = t & 7, blk = t >> 3;            // 32 blocks per CTA: 8 wide, 4 tall
  const int bx = blockIdx.x * 8 + (blk & 7), by = blockIdx.y * 4 + (blk >> 3);
  const bool inside = bx < g.w8 && by < g.h8;
  unsigned sa = 0, sb = 0, saa = 0, sbb = 0, sab = 0;
  if (inside) {
    const size_t o = (size_t)(by * 8 + row) * g.stride[0] + bx * 8;
    const uint4 va = *reinterpret_cast<const uint4*>(a0 + o), vb = *reinterpret_cast<const uint4*>(b0 + o);
    const unsigned wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const unsigned a_lo = wa[k] & 0xFFFF, a_hi = wa[k] >> 16, b_lo = wb[k] & 0xFFFF, b_hi = wb[k] >> 16;
      sa += a_lo + a_hi; sb += b_lo + b_hi;
      saa += a_lo * a_lo + a_hi * a_hi; sbb += b_lo * b_lo + b_hi * b_hi; sab += a_lo * b_lo + a_hi * b_hi;
    }
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o);
    saa += __shfl_xor_sync(0xffffffffu, saa, o); sbb += __shfl_xor_sync(0xffffffffu, sbb, o); sab += __shfl_xor_sync(0xffffffffu, sab, o);
  }
  unsigned long long sse = 0;
  float ssim = 0.f;
  unsigned cnt = 0;
  if (inside && row == 0) {
    sse = (unsigned long long)saa + sbb - 2ull * sab;
    const float L = (float)((1 << bit_depth) - 1), c1 = 0.0001f * L * L, c2 = 0.0009f * L * L;
    const float ma = sa * (1.f / 64), mb = sb * (1.f / 64);
    const float va = saa * (1.f / 64) - ma * ma, vb = sbb * (1.f / 64) - mb * mb, cov = sab * (1.f / 64) - ma * mb;
    ssim = ((2 * ma * mb + c1) * (2 * cov + c2)) / ((ma * ma + mb * mb + c1) * (va + vb + c2));
    cnt = 1;
  }
#pragma unroll
  for (int o = 8; o < 32; o <<= 1) {
    sse += __shfl_xor_sync(0xffffffffu, sse, o); ssim += __shfl_xor_sync(0xffffffffu, ssim, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((t & 31) == 0 && cnt) {
    atomicAdd(&out[frame].sse, sse);
    atomicAdd(&out[frame].ssim_sum, ssim);
    atomicAdd(&out[frame].blocks, cnt);
  }
}

}  // namespace

cudaError_t launch_quality(const Av1bGeom& g, int bit_depth, const uint16_t* rec_y, const uint16_t* src_y, size_t plane_elems,
                           QualityAcc* out, int n_frames, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(QualityAcc) * (size_t)n_frames, s);
  if (e != cudaSuccess) return e;
  dim3 grid((g.w8 + 7) / 8, (g.h8 + 3) / 4, n_frames);
  quality_kernel<<<grid, 256, 0, s>>>(g, bit_depth, rec_y, src_y, plane_elems, out);
  return cudaGetLastError();
}

}  // namespace av1b
