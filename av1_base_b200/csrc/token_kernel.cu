// Device tokenizer for inter frames: turns the block side information and the packed coefficient symbols of a
// batch of frames into the per-tile token lists the host range coder walks (tokens.h has the format and the
// per-block derivation, shared with its CPU statement).
//
// Decomposition: nothing here depends on entropy-coder state, so every block is independent once the mode
// classes (NEWMV or not, which the neighbours' contexts read) are known:
//   tok_mode_kernel   one thread per 8x8 unit that is a block origin: motion vector stack -> mode class
//   tok_walk_kernel   one warp per superblock (in coding order: tile by tile), one lane per block (Z order);
//                     <false>: token counts per block and per superblock, <true>: tokens at their final offsets
//   tok_scan_kernel   exclusive scan of the superblock counts (one CTA), tiles become contiguous ranges
// Replaces host work behind /root/reference/crates/daemon/src/encode/av1an.rs:126-139.
#include <cuda_runtime.h>
#include <stdint.h>
#include "av1_tables_dev.cuh"
#include "kernels.cuh"
#include "tokens.h"

namespace av1b {
namespace {

__device__ __forceinline__ TokFrame frame_view(const TokLaunch& P, int f) {
  TokFrame F;
  const size_t nsb = (size_t)P.g.sb_rows * P.g.sb_cols;
  F.blocks = P.blocks + (size_t)f * P.map_elems;
  F.mode_cls = P.mode_cls + (size_t)f * P.map_elems;
  for (int p = 0; p < 3; p++) { F.digest[p] = P.digest[p] + (size_t)f * P.plane_elems[p]; F.coef[p] = P.coef[p] + (size_t)f * P.plane_elems[p]; }
  F.cdef_idx = P.cdef_idx + (size_t)f * nsb;
  F.w8 = P.g.w8; F.h8 = P.g.h8; F.mi_cols = P.g.mi_cols; F.mi_rows = P.g.mi_rows; F.sb_cols = P.g.sb_cols;
  F.cdef_bits = ((P.nocdef_mask >> f) & 1) ? 0 : P.cdef_bits;
  F.scan[0] = tbl::scan_default_4; F.scan[1] = tbl::scan_default_8; F.scan[2] = tbl::scan_default_16;
  F.lr_units = P.lr_units ? P.lr_units + (size_t)f * P.lr_rows * P.lr_cols : nullptr;
  F.lr_rows = P.lr_rows; F.lr_cols = P.lr_cols;
  F.tx_sym_16 = 3; F.tx_sym_8 = 7;   // av1t_ext_tx_ind[4][DCT_DCT], av1t_ext_tx_ind[5][DCT_DCT] (checked on the host at launch)
  return F;
}

__device__ __forceinline__ TokTile tile_view(const TokLaunch& P, int tile) {
  const int tr = tile / P.g.tile_cols, tc = tile % P.g.tile_cols;
  TokTile T;
  T.mi_row_start = P.g.tile_row_start_sb[tr] * 16; T.mi_row_end = min(P.g.tile_row_start_sb[tr + 1] * 16, P.g.mi_rows);
  T.mi_col_start = P.g.tile_col_start_sb[tc] * 16; T.mi_col_end = min(P.g.tile_col_start_sb[tc + 1] * 16, P.g.mi_cols);
  return T;
}

__global__ void __launch_bounds__(128) tok_mode_kernel(const __grid_constant__ TokLaunch P) {
  const int f = blockIdx.y;
  if (!((P.inter_mask >> f) & 1)) return;
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= P.g.w8 * P.g.h8) return;
  const int ux = u % P.g.w8, uy = u / P.g.w8;
  const TokFrame F = frame_view(P, f);
  const Av1bBlockInfo& b = F.blocks[u];
  const int n8 = 1 << (b.blk_log2 - 3);
  if ((ux | uy) & (n8 - 1)) return;
  const TokTile T = tile_view(P, P.tile_of_sb[(uy >> 3) * P.g.sb_cols + (ux >> 3)]);
  const uint8_t cls = (uint8_t)tok_mode_class(F, T, uy * 2, ux * 2, b.blk_log2);
  uint8_t* mc = P.mode_cls + (size_t)f * P.map_elems;
  for (int yy = 0; yy < n8 && uy + yy < P.g.h8; yy++)
    for (int xx = 0; xx < n8 && ux + xx < P.g.w8; xx++) mc[(size_t)(uy + yy) * P.g.w8 + ux + xx] = cls;
}

// position of the (j+1)-th set bit of a 64-bit mask given as two halves
__device__ __forceinline__ int nth_set(unsigned lo, unsigned hi, int j) {
  const int nlo = __popc(lo);
  return j < nlo ? (int)__fns(lo, 0, j + 1) : 32 + (int)__fns(hi, 0, j - nlo + 1);
}

template <bool kEmit>
__global__ void __launch_bounds__(128) tok_walk_kernel(const __grid_constant__ TokLaunch P) {
  const int f = blockIdx.y, lane = threadIdx.x & 31;
  const int nsb = P.g.sb_rows * P.g.sb_cols;
  const int k = blockIdx.x * 4 + (threadIdx.x >> 5);   // coding-order index of this warp's superblock
  if (k >= nsb) return;
  if (!((P.inter_mask >> f) & 1)) {
    if (!kEmit && lane == 0) P.sb_off[(size_t)f * nsb + k] = 0;
    return;
  }
  const int sb = (int)P.sb_of_order[k];
  const int sbx = sb % P.g.sb_cols, sby = sb / P.g.sb_cols;
  const TokFrame F = frame_view(P, f);
  const TokTile T = tile_view(P, P.tile_of_sb[sb]);
  // block origins of the superblock in Z order: bit m of (lo, hi) = Morton unit m starts a block
  unsigned org[2], nskip[2];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int m = lane + 32 * h;
    const int ux = (m & 1) | ((m >> 1) & 2) | ((m >> 2) & 4), uy = ((m >> 1) & 1) | ((m >> 2) & 2) | ((m >> 3) & 4);
    const int gx = sbx * 8 + ux, gy = sby * 8 + uy;
    bool o = false, ns = false;
    if (gx < P.g.w8 && gy < P.g.h8) {
      const Av1bBlockInfo& b = F.blocks[(size_t)gy * P.g.w8 + gx];
      const int n8 = 1 << (b.blk_log2 - 3);
      o = ((ux | uy) & (n8 - 1)) == 0;
      ns = o && !b.skip;
    }
    org[h] = __ballot_sync(0xffffffffu, o);
    nskip[h] = __ballot_sync(0xffffffffu, ns);
  }
  const int nblk = __popc(org[0]) + __popc(org[1]);
  const int first_ns = nskip[0] ? __ffs(nskip[0]) - 1 : (nskip[1] ? 32 + __ffs(nskip[1]) - 1 : -1);   // carries cdef_idx
  uint32_t* bc = P.blk_count + (size_t)f * P.map_elems;
  uint32_t base = kEmit ? P.sb_off[(size_t)f * nsb + k] : 0;
  uint32_t sb_total = 0;
  {
    // restoration unit parameters precede the superblock's blocks
    uint32_t n_pre = 0;
    if (lane == 0) {
      TokSink K{kEmit ? P.tokens + base : nullptr, 0, (kEmit && base < P.cap) ? P.cap - base : 0};
      tok_sb_lr(F, sby * 16, sbx * 16, K);
      n_pre = K.n;
    }
    n_pre = __shfl_sync(0xffffffffu, n_pre, 0);
    base += n_pre; sb_total += n_pre;
  }
  for (int j0 = 0; j0 < nblk; j0 += 32) {
    const int j = j0 + lane;
    const bool act = j < nblk;
    int m = 0, r = 0, c = 0;
    size_t unit = 0;
    if (act) {
      m = nth_set(org[0], org[1], j);
      const int ux = (m & 1) | ((m >> 1) & 2) | ((m >> 2) & 4), uy = ((m >> 1) & 1) | ((m >> 2) & 2) | ((m >> 3) & 4);
      r = (sby * 8 + uy) * 2; c = (sbx * 8 + ux) * 2;
      unit = (size_t)(r >> 1) * P.g.w8 + (c >> 1);
    }
    if (!kEmit) {
      uint32_t n = 0;
      if (act) {
        TokSink K{nullptr, 0, 0};
        tok_block(F, T, r, c, m == first_ns, K);
        n = K.n;
        bc[unit] = n;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
      sb_total += n;
    } else {
      const uint32_t n = act ? bc[unit] : 0;
      uint32_t incl = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const uint32_t off = base + incl - n;
      if (act) {
        TokSink K{P.tokens + off, 0, off < P.cap ? P.cap - off : 0};
        tok_block(F, T, r, c, m == first_ns, K);
      }
      base += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  if (!kEmit && lane == 0) P.sb_off[(size_t)f * nsb + k] = sb_total;
}

// in-place exclusive scan of v[0..n) by one CTA; v[n] = total
__global__ void __launch_bounds__(1024) tok_scan_kernel(uint32_t* v, int n) {
  __shared__ uint32_t part[1024];
  const int t = threadIdx.x, per = (n + 1023) / 1024;
  const int i0 = min(t * per, n), i1 = min(i0 + per, n);
  uint32_t s = 0;
  for (int i = i0; i < i1; i++) s += v[i];
  part[t] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const uint32_t a = t >= o ? part[t - o] : 0;
    __syncthreads();
    part[t] += a;
    __syncthreads();
  }
  uint32_t run = part[t] - s;
  for (int i = i0; i < i1; i++) { const uint32_t x = v[i]; v[i] = run; run += x; }
  if (t == 1023) v[n] = part[1023];
}

}  // namespace

cudaError_t launch_scan_u32(uint32_t* v, int n, cudaStream_t s) {
  tok_scan_kernel<<<1, 1024, 0, s>>>(v, n);
  return cudaGetLastError();
}

cudaError_t launch_tok_count(const TokLaunch& p, cudaStream_t s) {
  const int units = p.g.w8 * p.g.h8, nsb = p.g.sb_rows * p.g.sb_cols;
  tok_mode_kernel<<<dim3((units + 127) / 128, p.n_frames), 128, 0, s>>>(p);
  tok_walk_kernel<false><<<dim3((nsb + 3) / 4, p.n_frames), 128, 0, s>>>(p);
  tok_scan_kernel<<<1, 1024, 0, s>>>(p.sb_off, p.n_frames * nsb);
  return cudaGetLastError();
}

cudaError_t launch_tok_emit(const TokLaunch& p, cudaStream_t s) {
  const int nsb = p.g.sb_rows * p.g.sb_cols;
  tok_walk_kernel<true><<<dim3((nsb + 3) / 4, p.n_frames), 128, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace av1b
