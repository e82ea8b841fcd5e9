// Range coding of the token lists on the device: one warp per (frame, tile).
//
// The arithmetic coder of a tile is a serial chain (range / low and the adapting CDFs), but tiles are
// independent: a batch of 8 4K inter frames has 840 of them (about 6x6 superblocks each), and with the token lists
// already on the device the chain costs a warp about 80 instructions (0.25 to 0.5 us) a symbol while the 148 SMs
// are busy with the next batches' kernels; the latency of the kernel is that of the tile with the most tokens
// (12 ms at 4K, hidden behind two to three batches).  That frees the host (32 cores for 8 GPUs on these boxes: with
// the coder on the host the box stops scaling at two GPUs) for headers and concatenation only.
//   rc_code_kernel    warp per tile: CDFs of the tile in shared memory (lane i adapts entry i), range coder state
//                     in registers (warp-uniform), bytes with in-place carry propagation into a per-tile region
//   tok_scan (reused) exclusive scan of the tile byte counts
//   rc_gather_kernel  warp per tile: region -> contiguous byte stream of the batch
// Bit-exact with RangeEncoder / pack_tile_tokens (bitstream.h / bitstream.cc), which stay the CPU statement.
// Replaces host work behind /root/reference/crates/daemon/src/encode/av1an.rs:126-139.
#include <cuda_runtime.h>
#include <stdint.h>
#include "av1_tables_dev.cuh"
#include "kernels.cuh"
#include "tokens.h"

namespace av1b {
namespace {

struct Rc {
  uint32_t low, rng;
  int cnt;
  uint8_t* buf;
  uint32_t n, cap;
};

// one output unit of the coder (8 data bits + carry), lane 0 stores; the carry goes into the bytes already written
__device__ __forceinline__ void rc_put(Rc& c, uint32_t v, int lane) {
  v &= 0xFFFFu;
  if (lane == 0 && c.n < c.cap) {
    uint32_t add = v >> 8, j = c.n;
    while (add && j > 0) { j--; const uint32_t sum = c.buf[j] + add; c.buf[j] = (uint8_t)sum; add = sum >> 8; }
    c.buf[c.n] = (uint8_t)v;
  }
  c.n++;
}

// RangeEncoder::encode
__device__ __forceinline__ void rc_encode(Rc& c, int s, uint32_t fh, uint32_t fl, int n, int lane) {
  uint32_t r = c.rng, l = c.low;
  const int N = n - 1;
  const uint32_t v = ((r >> 8) * (fh >> 6) >> 1) + 4 * (N - s);
  const uint32_t uc = ((r >> 8) * (fl >> 6) >> 1) + 4 * (N - s + 1);
  const uint32_t u = s > 0 ? uc : r;
  l += r - u;
  r = u - v;
  const int d = __clz(r) - 16;
  int cc = c.cnt;
  int sft = cc + d;
  if (sft >= 0) {
    cc += 16;
    uint32_t m = (1u << cc) - 1;
    if (sft >= 8) {
      rc_put(c, l >> cc, lane);
      l &= m;
      cc -= 8;
      m >>= 8;
    }
    rc_put(c, l >> cc, lane);
    sft = cc + d - 24;
    l &= m;
  }
  c.low = l << d;
  c.rng = r << d;
  c.cnt = sft;
}

__device__ __forceinline__ void rc_bool(Rc& c, int b, int lane) { rc_encode(c, b, b ? 0u : 16384u, 16384u, 2, lane); }
__device__ __forceinline__ void rc_literal(Rc& c, uint32_t v, int nbits, int lane) {
  for (int i = nbits - 1; i >= 0; i--) rc_bool(c, (v >> i) & 1, lane);
}

// RangeEncoder::symbol: icdf in shared memory.  Lane i <= n loads entry i once (entry n is the adaptation counter);
// the two CDF values the interval needs and the counter travel by shuffle, so the only barrier is the one that
// orders this symbol's stores before the next symbol's loads.
__device__ __forceinline__ void rc_symbol(Rc& c, int s, uint16_t* icdf, int n, int lane) {
  const int x = lane <= n ? icdf[lane] : 0;
  const uint32_t fh = (uint32_t)__shfl_sync(0xffffffffu, x, s), fl = (uint32_t)__shfl_sync(0xffffffffu, x, s - (s > 0));
  const int cntv = __shfl_sync(0xffffffffu, x, n);
  const int rate = 3 + (cntv > 15) + (cntv > 31) + (n > 3 ? 2 : 1);
  if (lane < n - 1) icdf[lane] = (uint16_t)(lane < s ? x + ((32768 - x) >> rate) : x - (x >> rate));
  else if (lane == n) icdf[n] = (uint16_t)(cntv + (cntv < 32));
  rc_encode(c, s, fh, fl, n, lane);
  __syncwarp();
}

// sub-exponential codes of the restoration coefficients (bitstream.cc lr_put_*)
__device__ void rc_uniform(Rc& c, int v, int n, int lane) {
  const int w = 32 - __clz((unsigned)n), m = (1 << w) - n;
  if (v < m) rc_literal(c, v, w - 1, lane);
  else { rc_literal(c, m + ((v - m) >> 1), w - 1, lane); rc_literal(c, (v - m) & 1, 1, lane); }
}
__device__ void rc_subexp(Rc& c, int x, int num_syms, int k, int lane) {
  int i = 0, mk = 0;
  for (;;) {
    const int b2 = i ? k + i - 1 : k, a = 1 << b2;
    if (num_syms <= mk + 3 * a) { rc_uniform(c, x - mk, num_syms - mk, lane); return; }
    const int more = x >= mk + a;
    rc_literal(c, more, 1, lane);
    if (more) { i++; mk += a; }
    else { rc_literal(c, x - mk, b2, lane); return; }
  }
}
__device__ __forceinline__ int rc_recenter(int r, int v) { return v > 2 * r ? v : (v >= r ? (v - r) << 1 : ((r - v) << 1) - 1); }
__device__ void rc_signed_subexp_with_ref(Rc& c, int v, int low, int high, int k, int r, int lane) {
  const int mx = high - low, vv = v - low, rr = r - low;
  const int x = (rr << 1) <= mx ? rc_recenter(rr, vv) : rc_recenter(mx - 1 - rr, mx - 1 - vv);
  rc_subexp(c, x, mx, k, lane);
}

__global__ void __launch_bounds__(32) rc_code_kernel(const __grid_constant__ RcLaunch P) {
  // only the CDFs inter frames code (the head of TileCdfs) live in shared memory
  constexpr int kWords = (int)(offsetof(TileCdfs, kf_y_mode) / 4);
  __shared__ uint32_t cdf_words[kWords];
  const int tile = blockIdx.x, f = blockIdx.y, lane = threadIdx.x;
  const int nsb = P.nsb, n_tiles = P.n_tiles;
  uint32_t* len_out = P.tile_len + (size_t)f * n_tiles + tile;
  if (!((P.inter_mask >> f) & 1)) { if (lane == 0) *len_out = 0; return; }
  const uint32_t* off = P.sb_off + (size_t)f * nsb;
  const uint32_t t0 = off[P.tile_first_k[tile]], t1 = off[P.tile_first_k[tile + 1]];
  if (t1 > P.tok_cap) { if (lane == 0) *len_out = 0; return; }   // token buffer overflowed: the host grows it and codes again
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(((P.alt_mask >> f) & 1) ? P.cdf_init_alt : P.cdf_init);
    for (int i = lane; i < kWords; i += 32) cdf_words[i] = src[i];
  }
  int ref_wiener[2][3] = {{3, -7, 15}, {3, -7, 15}}, ref_sgr[2] = {-32, 31};   // luma plane only
  __syncwarp();
  uint16_t* base = reinterpret_cast<uint16_t*>(cdf_words);
  Rc c;
  c.low = 0; c.rng = 0x8000; c.cnt = -9; c.n = 0;
  // a symbol emits at most two bytes (15 bits): the region of a tile holds 2 bytes per token + 64 for the
  // terminating bytes and the rare longer Golomb escape; writes beyond it are dropped and reported
  c.buf = P.region + (size_t)2 * t0 + (size_t)64 * ((size_t)f * n_tiles + tile);
  c.cap = 2 * (t1 - t0) + 64;
  for (uint32_t i0 = t0; i0 < t1; i0 += 32) {
    const uint32_t mine = i0 + lane < t1 ? P.tokens[i0 + lane] : 0;
    const int cntk = (int)min(32u, t1 - i0);
    for (int k = 0; k < cntk; k++) {
      const uint32_t t = __shfl_sync(0xffffffffu, mine, k), o = t & 0xFFFFu;
      if (o < TOK_FIRST_SPECIAL) {
        rc_symbol(c, (int)(t >> 21), base + o, (int)((t >> 16) & 31), lane);
      } else if (o == TOK_RAW) {
        rc_literal(c, t >> 21, (int)((t >> 16) & 31), lane);
      } else if (o == TOK_GOLOMB) {
        const uint32_t x = t >> 16;
        const int len = 32 - __clz(x);
        for (int q = 0; q < len - 1; q++) rc_bool(c, 0, lane);
        for (int q = len - 1; q >= 0; q--) rc_bool(c, (x >> q) & 1, lane);
      } else if (o == TOK_LR) {
        const int kind = (t >> 16) & 1, pass = (t >> 19) & 1, j = (t >> 20) & 15, v = (int)(t >> 24) - 128;
        if (kind == 0) {
          const int tmin = j == 0 ? -5 : j == 1 ? -23 : -17, tmax = j == 0 ? 10 : j == 1 ? 8 : 46;
          rc_signed_subexp_with_ref(c, v, tmin, tmax + 1, j + 1, ref_wiener[pass][j], lane);
          ref_wiener[pass][j] = v;
        } else {
          const int xmin = pass ? -32 : -96, xmax = pass ? 95 : 31;
          if (tbl::sgr_params[j][pass]) {
            rc_signed_subexp_with_ref(c, v, xmin, xmax + 1, 4, ref_sgr[pass], lane);
            ref_sgr[pass] = v;
          } else {
            int w = 0;
            if (pass == 1) w = min(max(128 - ref_sgr[0], -32), 95);
            ref_sgr[pass] = w;
          }
        }
      } else {
        // forced split at the picture edge: probability gathered from the adaptive partition CDF, no adaptation
        const uint16_t* pc = base + AV1B_CDF_OFF(partition) + ((t >> 16) & 31) * 11;
        const bool has_cols = (t >> 21) & 1, is8 = (t >> 22) & 1;
        auto prob = [&](int q) -> int { return (q > 0 ? (int)pc[q - 1] : 32768) - (int)pc[q]; };
        int psum;
        if (has_cols) { psum = prob(2) + prob(3); if (!is8) psum += prob(4) + prob(6) + prob(7) + prob(9); }
        else { psum = prob(1) + prob(3); if (!is8) psum += prob(4) + prob(5) + prob(6) + prob(8); }
        rc_encode(c, 1, 0u, (uint32_t)(psum & 0xFFFF), 2, lane);
      }
    }
  }
  // RangeEncoder::finish
  {
    uint32_t l = c.low;
    int cc = c.cnt, s = 10;
    const uint32_t m = 0x3FFF;
    uint32_t e = ((l + m) & ~m) | (m + 1);
    s += cc;
    if (s > 0) {
      uint32_t nmask = (1u << (cc + 16)) - 1;
      do {
        rc_put(c, e >> (cc + 16), lane);
        e &= nmask;
        s -= 8;
        cc -= 8;
        nmask >>= 8;
      } while (s > 0);
    }
  }
  if (lane == 0) {
    *len_out = min(c.n, c.cap);
    if (c.n > c.cap) atomicOr(P.overflow, 1u);
  }
}

__global__ void __launch_bounds__(128) rc_gather_kernel(const __grid_constant__ RcLaunch P) {
  const int tile = blockIdx.x, f = blockIdx.y;
  const size_t idx = (size_t)f * P.n_tiles + tile;
  const uint32_t o0 = P.tile_len[idx], o1 = P.tile_len[idx + 1];   // after the scan: offsets
  if (o1 == o0) return;
  const uint32_t t0 = P.sb_off[(size_t)f * P.nsb + P.tile_first_k[tile]];
  if (t0 > P.tok_cap) return;
  const uint8_t* src = P.region + (size_t)2 * t0 + (size_t)64 * idx;
  for (uint32_t i = threadIdx.x; i < o1 - o0; i += blockDim.x)
    if (o0 + i < P.cap_bytes) P.bytes[o0 + i] = src[i];
}

}  // namespace

cudaError_t launch_rc(const RcLaunch& p, cudaStream_t s) {
  rc_code_kernel<<<dim3(p.n_tiles, p.n_frames), 32, 0, s>>>(p);
  cudaError_t e = launch_scan_u32(p.tile_len, p.n_tiles * p.n_frames, s);
  if (e != cudaSuccess) return e;
  rc_gather_kernel<<<dim3(p.n_tiles, p.n_frames), 128, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace av1b
