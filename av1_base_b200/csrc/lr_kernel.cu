// AV1 loop restoration (spec 7.17) for sm_100a: separable 7-tap Wiener filter and the self-guided
// (box r=2 / r=1) filter.  One CTA owns a 64-column (32 for chroma) slice of one 64-row stripe
// (stripes are offset by 8 luma rows, spec 7.17.? "StripeStartY"), so that the restoration unit and
// the stripe-boundary rule are uniform for the CTA: rows outside the stripe come from the DEBLOCKED
// (pre-CDEF) frame, at most two rows deep, rows inside from the CDEF output.  The window (+3 halo)
// is staged once in shared memory; Wiener keeps its horizontal intermediate as int16 in shared
// memory, the self-guided filter builds its A/B planes there.  Frame read once (+ boundary rows of
// the deblocked frame), written once: algorithmic bytes 2.125*S.
//
// Replaces arithmetic the reference delegates to av1an + SVT-AV1
// (/root/reference/crates/daemon/src/encode/av1an.rs:126-139).
// Bit-exact against oracle/av1_oracle.cpp orc_lr_frame (pinned through dav1d and libaom).
#include <cuda_runtime.h>
#include <stdint.h>
#include "av1_tables_dev.cuh"
#include "kernels.cuh"

namespace av1b {
namespace {

constexpr int kThreads = 256;
constexpr int kH = 3;                       // halo
constexpr int kWinW = 64 + 2 * kH;          // 70
constexpr int kWinStride = kWinW + 1;       // 71
constexpr int kABW = 66;                    // A/B planes cover rows/cols -1 .. 64

struct Smem {
  uint16_t win[kWinW * kWinStride];         // rows ys-3 .. ye+2
  union {
    int16_t inter[kWinW * 64];              // Wiener: horizontal pass, rows ys-3 .. ye+2
    struct { uint16_t A[kABW * kABW]; int32_t B[kABW * kABW]; } ab;
  } u;
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// A/B of the box filter with radius r at window position (i, j) (tile coordinates, -1 .. h / w)
__device__ __forceinline__ void sgr_ab(const Smem& sm, int i, int j, int r, int s, int bd, int* a_out, int* b_out) {
  uint32_t a = 0, b = 0;
  for (int dy = -r; dy <= r; dy++)
    for (int dx = -r; dx <= r; dx++) {
      const uint32_t c = sm.win[(i + kH + dy) * kWinStride + j + kH + dx];
      a += c * c; b += c;
    }
  const int n = (2 * r + 1) * (2 * r + 1);
  const uint32_t one_by_n = (uint32_t)tbl::one_by_x[n - 1];
  const uint32_t a_r = bd > 8 ? (a + (1u << (2 * (bd - 8) - 1))) >> (2 * (bd - 8)) : a;
  const uint32_t d = bd > 8 ? (b + (1u << (bd - 8 - 1))) >> (bd - 8) : b;
  const uint32_t p = (a_r * n < d * d) ? 0 : a_r * n - d * d;
  const uint32_t z = (uint32_t)(((uint64_t)p * (uint32_t)s + (1u << 19)) >> 20);
  const uint32_t a2 = (uint32_t)tbl::x_by_xplus1[z < 255u ? z : 255u];
  const uint32_t b2 = (256 - a2) * b * one_by_n;
  *a_out = (int)a2;
  *b_out = (int)((b2 + (1u << 11)) >> 12);
}

// kSearch: no output; the luma plane is filtered with the Wiener and the self-guided parameters of P.cand and the
// squared errors against the source (and that of the unfiltered CDEF output) are added to P.sse[3][units].
template <bool kSearch>
__global__ void __launch_bounds__(kThreads) lr_kernel(const LrLaunch P) {
  __shared__ Smem sm;
  const Av1bGeom& g = P.g;
  const int tid = threadIdx.x;
  const int tx = blockIdx.x, stripe = blockIdx.y, frame = blockIdx.z;
  const int bd = P.bit_depth, maxv = (1 << bd) - 1;
  unsigned acc[3] = {0, 0, 0};   // search: squared errors of this thread (none, Wiener, self-guided)
  int s_urow = 0, s_ucol = 0;
  for (int p = 0; p < (kSearch ? 1 : 3); p++) {
    const int ss = p > 0;
    const int TW = 64 >> ss;
    const int pw = (g.width + ss) >> ss, ph = (g.height + ss) >> ss;
    const int stride = g.stride[p];
    const int x0 = tx * TW;
    const int s_start = (-8 + stripe * 64) >> ss, s_end = s_start + (64 >> ss) - 1;
    const int ys = max(s_start, 0), ye = min(s_end + 1, ph);   // rows [ys, ye)
    const int h = ye - ys;
    const uint16_t* cdef = P.cdef[p] + (size_t)frame * P.plane_elems[p];
    const uint16_t* deb = P.deb[p] + (size_t)frame * P.plane_elems[p];
    uint16_t* out = P.out[p] + (size_t)frame * P.plane_elems[p];
    __syncthreads();
    if (h <= 0 || x0 >= stride) continue;
    int type = AV1B_RESTORE_NONE;
    Av1bLrUnit unit;
    const uint16_t* srcp = nullptr;
    if (kSearch) {
      if (x0 >= pw) continue;
      unit = P.cand;
      srcp = P.src_y + (size_t)frame * P.plane_elems[0];
      s_urow = min(P.unit_rows[0] - 1, (ys + 8) / P.unit_size[0]);
      s_ucol = min(P.unit_cols[0] - 1, x0 / P.unit_size[0]);
    } else if (P.lr_type[p] != AV1B_RESTORE_NONE && P.units[p]) {
      const int us = P.unit_size[p];
      const int urow = min(P.unit_rows[p] - 1, (((ys << ss) + 8) >> ss) / us);
      const int ucol = min(P.unit_cols[p] - 1, x0 / us);
      unit = P.units[p][(size_t)frame * P.unit_rows[p] * P.unit_cols[p] + urow * P.unit_cols[p] + ucol];
      type = x0 < pw ? unit.type : AV1B_RESTORE_NONE;
    }
    if (!kSearch && type == AV1B_RESTORE_NONE) {
      for (int o = tid; o < h * (TW / 8); o += kThreads) {
        const int r = o / (TW / 8), v = o % (TW / 8);
        const size_t off = (size_t)(ys + r) * stride + x0 + v * 8;
        *reinterpret_cast<uint4*>(out + off) = *reinterpret_cast<const uint4*>(cdef + off);
      }
      continue;
    }
    // ---- stage the window with the stripe-boundary rule ----
    const int wrows = h + 2 * kH, wcols = TW + 2 * kH;
    for (int o = tid; o < wrows * wcols; o += kThreads) {
      const int r = o / wcols, c = o % wcols;
      const int x = clampi(x0 - kH + c, 0, pw - 1);
      int y = clampi(ys - kH + r, 0, ph - 1);
      const uint16_t* srcp = cdef;
      if (y < s_start) { y = max(s_start - 2, y); srcp = deb; }
      else if (y > s_end) { y = min(s_end + 2, y); srcp = deb; }
      sm.win[r * kWinStride + c] = srcp[(size_t)y * stride + x];
    }
    __syncthreads();
    for (int pass = 0; pass < (kSearch ? 2 : 1); pass++) {
    if (kSearch) { type = pass ? AV1B_RESTORE_SGRPROJ : AV1B_RESTORE_WIENER; if (pass) __syncthreads(); }
    if (type == AV1B_RESTORE_WIENER) {
      int hf[7], vf[7];
      hf[3] = vf[3] = 128;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        hf[i] = hf[6 - i] = unit.wiener_h[i]; hf[3] -= 2 * unit.wiener_h[i];
        vf[i] = vf[6 - i] = unit.wiener_v[i]; vf[3] -= 2 * unit.wiener_v[i];
      }
      const int offset = 1 << (bd + 3), limit = (1 << (bd + 5)) - 1;
      for (int o = tid; o < wrows * TW; o += kThreads) {
        const int r = o / TW, c = o % TW;
        const uint16_t* w = sm.win + r * kWinStride + c;
        int s = 0;
#pragma unroll
        for (int t = 0; t < 7; t++) s += hf[t] * (int)w[t];
        sm.u.inter[r * 64 + c] = (int16_t)clampi((s + 4) >> 3, -offset, limit - offset);
      }
      __syncthreads();
      for (int o = tid; o < h * TW; o += kThreads) {
        const int r = o / TW, c = o % TW;
        int v;
        if (x0 + c < pw) {
          int s = 0;
#pragma unroll
          for (int t = 0; t < 7; t++) s += vf[t] * (int)sm.u.inter[(r + t) * 64 + c];
          v = clampi((s + (1 << 10)) >> 11, 0, maxv);
        } else {
          v = sm.win[(r + kH) * kWinStride + c + kH];
        }
        if (kSearch) {
          if (x0 + c < pw) { const int d = v - (int)srcp[(size_t)(ys + r) * stride + x0 + c]; acc[1] += (unsigned)(d * d); }
        } else {
          out[(size_t)(ys + r) * stride + x0 + c] = (uint16_t)v;
        }
      }
    } else {
      const int set = unit.sgr_set;
      const int r0 = tbl::sgr_params[set][0], r1 = tbl::sgr_params[set][1];
      const int s0 = tbl::sgr_params[set][2], s1 = tbl::sgr_params[set][3];
      const int w0 = unit.sgr_xqd[0], w1 = unit.sgr_xqd[1], w2 = 128 - w0 - w1;
      const int per_thread = (64 * 64) / kThreads;   // 16 (luma), 4 used for chroma
      int flt0[per_thread];
      const int npx = h * TW;
      if (r0) {
        // A/B only on odd absolute rows (the r=2 pass subsamples rows)
        for (int o = tid; o < (h + 2) * (TW + 2); o += kThreads) {
          const int i = o / (TW + 2) - 1, j = o % (TW + 2) - 1;
          if (!((ys + i) & 1)) continue;
          int a, b;
          sgr_ab(sm, i, j, r0, s0, bd, &a, &b);
          sm.u.ab.A[(i + 1) * kABW + j + 1] = (uint16_t)a;
          sm.u.ab.B[(i + 1) * kABW + j + 1] = b;
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < per_thread; it++) {
          const int q = it * kThreads + tid;
          flt0[it] = 0;
          if (q < npx) {
            const int i = q / TW, j = q % TW;
            const int src = sm.win[(i + kH) * kWinStride + j + kH];
            const uint16_t* Ap = sm.u.ab.A + (i + 1) * kABW + j + 1;
            const int32_t* Bp = sm.u.ab.B + (i + 1) * kABW + j + 1;
            int a, b, shift;
            if ((ys + i) & 1) {
              shift = 4;
              a = 6 * Ap[0] + 5 * (Ap[-1] + Ap[1]);
              b = 6 * Bp[0] + 5 * (Bp[-1] + Bp[1]);
            } else {
              shift = 5;
              a = 6 * (Ap[-kABW] + Ap[kABW]) + 5 * (Ap[-kABW - 1] + Ap[-kABW + 1] + Ap[kABW - 1] + Ap[kABW + 1]);
              b = 6 * (Bp[-kABW] + Bp[kABW]) + 5 * (Bp[-kABW - 1] + Bp[-kABW + 1] + Bp[kABW - 1] + Bp[kABW + 1]);
            }
            const int v = a * src + b;
            flt0[it] = (v + (1 << (8 + shift - 4 - 1))) >> (8 + shift - 4);
          }
        }
        __syncthreads();
      }
      if (r1) {
        for (int o = tid; o < (h + 2) * (TW + 2); o += kThreads) {
          const int i = o / (TW + 2) - 1, j = o % (TW + 2) - 1;
          int a, b;
          sgr_ab(sm, i, j, r1, s1, bd, &a, &b);
          sm.u.ab.A[(i + 1) * kABW + j + 1] = (uint16_t)a;
          sm.u.ab.B[(i + 1) * kABW + j + 1] = b;
        }
        __syncthreads();
      }
#pragma unroll
      for (int it = 0; it < per_thread; it++) {
        const int q = it * kThreads + tid;
        if (q < npx) {
          const int i = q / TW, j = q % TW;
          const int src = sm.win[(i + kH) * kWinStride + j + kH];
          int v = src;
          if (x0 + j < pw) {
            const int uu = src << 4;
            const int f0 = r0 ? flt0[it] : uu;
            int f1 = uu;
            if (r1) {
              const uint16_t* Ap = sm.u.ab.A + (i + 1) * kABW + j + 1;
              const int32_t* Bp = sm.u.ab.B + (i + 1) * kABW + j + 1;
              const int a = 4 * (Ap[0] + Ap[-1] + Ap[1] + Ap[-kABW] + Ap[kABW]) +
                            3 * (Ap[-kABW - 1] + Ap[-kABW + 1] + Ap[kABW - 1] + Ap[kABW + 1]);
              const int b = 4 * (Bp[0] + Bp[-1] + Bp[1] + Bp[-kABW] + Bp[kABW]) +
                            3 * (Bp[-kABW - 1] + Bp[-kABW + 1] + Bp[kABW - 1] + Bp[kABW + 1]);
              f1 = (a * src + b + (1 << 8)) >> 9;
            }
            v = clampi((w1 * uu + w0 * f0 + w2 * f1 + (1 << 10)) >> 11, 0, maxv);
          }
          if (kSearch) {
            if (x0 + j < pw) { const int d = v - (int)srcp[(size_t)(ys + i) * stride + x0 + j]; acc[2] += (unsigned)(d * d); }
          } else {
            out[(size_t)(ys + i) * stride + x0 + j] = (uint16_t)v;
          }
        }
      }
    }
    }   // pass
    if (kSearch) {
      for (int o = tid; o < h * TW; o += kThreads) {
        const int r = o / TW, c = o % TW;
        if (x0 + c < pw) {
          const int d = (int)sm.win[(r + kH) * kWinStride + c + kH] - (int)srcp[(size_t)(ys + r) * stride + x0 + c];
          acc[0] += (unsigned)(d * d);
        }
      }
    }
  }
  if (kSearch) {
    const size_t n_units = (size_t)P.unit_rows[0] * P.unit_cols[0];
    unsigned long long* sse = P.sse + (size_t)frame * 3 * n_units + (size_t)s_urow * P.unit_cols[0] + s_ucol;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      unsigned a = acc[k];
      for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if ((tid & 31) == 0 && a) atomicAdd(sse + k * n_units, (unsigned long long)a);
    }
  }
}

// one thread per restoration unit: NONE unless a candidate beats it by more than the bias; ties keep the earlier
// of NONE, WIENER, SGRPROJ (orc_lr_search)
__global__ void lr_decide_kernel(const unsigned long long* sse, int n_units, int n_frames, unsigned long long bias, Av1bLrUnit cand,
                                 Av1bLrUnit* units) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_units * n_frames) return;
  const int f = i / n_units, u = i % n_units;
  const unsigned long long* s = sse + (size_t)f * 3 * n_units + u;
  int best = 0;
  unsigned long long bv = s[0];
  for (int k = 1; k < 3; k++) {
    const unsigned long long v = s[(size_t)k * n_units] + bias;
    if (v < bv) { bv = v; best = k; }
  }
  Av1bLrUnit o = cand;
  o.type = (int8_t)(best == 0 ? AV1B_RESTORE_NONE : best == 1 ? AV1B_RESTORE_WIENER : AV1B_RESTORE_SGRPROJ);
  units[i] = o;
}

}  // namespace

cudaError_t launch_lr(const LrLaunch& p, int n_frames, cudaStream_t s) {
  const int stripes = (p.g.height + 8 + 63) / 64;
  dim3 grid(p.g.sb_cols, stripes, n_frames);
  lr_kernel<false><<<grid, kThreads, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_lr_search(const LrLaunch& p, int n_frames, long long bias, Av1bLrUnit* units_out, cudaStream_t s) {
  const int stripes = (p.g.height + 8 + 63) / 64;
  const int n_units = p.unit_rows[0] * p.unit_cols[0];
  cudaError_t e = cudaMemsetAsync(p.sse, 0, (size_t)n_frames * 3 * n_units * sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  dim3 grid(p.g.sb_cols, stripes, n_frames);
  lr_kernel<true><<<grid, kThreads, 0, s>>>(p);
  lr_decide_kernel<<<(n_units * n_frames + 127) / 128, 128, 0, s>>>(p.sse, n_units, n_frames, (unsigned long long)bias, p.cand, units_out);
  return cudaGetLastError();
}

}  // namespace av1b
