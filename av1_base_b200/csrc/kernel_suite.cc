// C-ABI entry points of the per-kernel suite (BASELINE.json config 2 "kernel bit-exact suite"):
// each call uploads host buffers, runs ONE CUDA kernel `reps` times between CUDA events on its own
// stream, and downloads the result of the last run.  Used by tests/test_gpu_kernel_suite.py (parity
// against the oracle and libaom's C functions) and by bench.py --kernels (per-kernel roofline).
// Boundary replaced: the codec arithmetic behind /root/reference/crates/daemon/src/encode/av1an.rs:126-139.
#include <cuda_runtime.h>
#include <string.h>
#include <vector>
#include "../../include/av1b200.h"
#include "capi_internal.h"
#include "kernels.cuh"
#include "av1_tables.h"

using namespace av1b;

#define CKS(call)                                                                    \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) {                                                         \
      set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));  \
      return AV1B_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

namespace {

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 1); }
  template <typename T> T* as() { return static_cast<T*>(p); }
};

struct Timer {
  cudaStream_t s = nullptr;
  cudaEvent_t a = nullptr, b = nullptr;
  ~Timer() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); if (s) cudaStreamDestroy(s); }
  cudaError_t init() {
    cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&a);
    if (e == cudaSuccess) e = cudaEventCreate(&b);
    return e;
  }
};

int select_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    set_error("no CUDA device visible (av1b200 has no CPU fallback)");
    return AV1B_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) { set_error("device %d out of range", device); return AV1B_ERR_INVALID; }
  if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return AV1B_ERR_CUDA; }
  return AV1B_OK;
}

// runs `launch` reps times (after one warm-up when reps > 1) and reports the mean duration
template <typename F>
int timed(Timer& t, int reps, double* ms, F launch) {
  if (reps < 1) reps = 1;
  if (reps > 1) { cudaError_t e = launch(); if (e != cudaSuccess) { set_error("kernel launch: %s", cudaGetErrorString(e)); return AV1B_ERR_CUDA; } }
  CKS(cudaEventRecord(t.a, t.s));
  for (int i = 0; i < reps; i++) {
    cudaError_t e = launch();
    if (e != cudaSuccess) { set_error("kernel launch: %s", cudaGetErrorString(e)); return AV1B_ERR_CUDA; }
  }
  CKS(cudaEventRecord(t.b, t.s));
  CKS(cudaStreamSynchronize(t.s));
  float f = 0;
  CKS(cudaEventElapsedTime(&f, t.a, t.b));
  if (ms) *ms = (double)f / reps;
  return AV1B_OK;
}

}  // namespace

extern "C" {

int av1b_k_inv_txfm_add(int device, const int32_t* coef, uint16_t* dst, int n_blocks, int w, int h, int tx_type,
                        int bit_depth, int reps, double* ms_per_launch) {
  if (!coef || !dst || n_blocks <= 0 || tx_type < 0 || tx_type > 15 || (bit_depth != 8 && bit_depth != 10)) {
    set_error("bad argument"); return AV1B_ERR_INVALID;
  }
  int rc = select_device(device);
  if (rc) return rc;
  Timer t; CKS(t.init());
  DevBuf dc, dd, d0;
  const size_t cb = (size_t)n_blocks * 1024 * 4, db = (size_t)n_blocks * w * h * 2;
  CKS(dc.alloc(cb)); CKS(dd.alloc(db)); CKS(d0.alloc(db));
  CKS(cudaMemcpyAsync(dc.p, coef, cb, cudaMemcpyHostToDevice, t.s));
  CKS(cudaMemcpyAsync(d0.p, dst, db, cudaMemcpyHostToDevice, t.s));
  rc = timed(t, reps, ms_per_launch, [&]() {
    // the kernel adds into dst: restore the prediction before every run (device-to-device, untimed share is small)
    cudaError_t e = cudaMemcpyAsync(dd.p, d0.p, db, cudaMemcpyDeviceToDevice, t.s);
    if (e != cudaSuccess) return e;
    return launch_inv_txfm_add(dc.as<int32_t>(), dd.as<uint16_t>(), n_blocks, w, h, tx_type, bit_depth, t.s);
  });
  if (rc) return rc;
  CKS(cudaMemcpyAsync(dst, dd.p, db, cudaMemcpyDeviceToHost, t.s));
  CKS(cudaStreamSynchronize(t.s));
  return AV1B_OK;
}

int av1b_k_fwd_txfm(int device, const int16_t* resid, int32_t* coef, int n_blocks, int w, int h, int tx_type, int reps,
                    double* ms_per_launch) {
  auto size_ok = [](int n) { return n == 4 || n == 8 || n == 16 || n == 32 || n == 64; };
  if (!resid || !coef || n_blocks <= 0 || tx_type < 0 || tx_type > 15 || !size_ok(w) || !size_ok(h) || w > 4 * h || h > 4 * w) {
    set_error("bad argument"); return AV1B_ERR_INVALID;
  }
  const int m = std::max(w, h);
  if ((m == 64 && tx_type != 0) || (m == 32 && tx_type != 0 && tx_type != 9)) { set_error("transform type not legal at this size"); return AV1B_ERR_INVALID; }
  int rc = select_device(device);
  if (rc) return rc;
  Timer t; CKS(t.init());
  DevBuf dr, dc;
  const size_t rb = (size_t)n_blocks * w * h * 2, cb = (size_t)n_blocks * std::min(w, 32) * std::min(h, 32) * 4;
  CKS(dr.alloc(rb)); CKS(dc.alloc(cb));
  CKS(cudaMemcpyAsync(dr.p, resid, rb, cudaMemcpyHostToDevice, t.s));
  rc = timed(t, reps, ms_per_launch, [&]() { return launch_fwd_txfm(dr.as<int16_t>(), dc.as<int32_t>(), n_blocks, w, h, tx_type, t.s); });
  if (rc) return rc;
  CKS(cudaMemcpyAsync(coef, dc.p, cb, cudaMemcpyDeviceToHost, t.s));
  CKS(cudaStreamSynchronize(t.s));
  return AV1B_OK;
}

struct FrameBufs {
  DevBuf d[3];
  size_t elems[3];
};

static int upload_planes(const Av1bGeom& g, int n_frames, const uint16_t* const h[3], FrameBufs& fb, cudaStream_t s) {
  for (int p = 0; p < 3; p++) {
    fb.elems[p] = (size_t)g.stride[p] * g.rows[p];
    CKS(fb.d[p].alloc(fb.elems[p] * n_frames * 2));
    if (h && h[p]) CKS(cudaMemcpyAsync(fb.d[p].p, h[p], fb.elems[p] * n_frames * 2, cudaMemcpyHostToDevice, s));
    else CKS(cudaMemsetAsync(fb.d[p].p, 0, fb.elems[p] * n_frames * 2, s));
  }
  return AV1B_OK;
}

static int download_planes(int n_frames, uint16_t* const h[3], FrameBufs& fb, cudaStream_t s) {
  for (int p = 0; p < 3; p++) CKS(cudaMemcpyAsync(h[p], fb.d[p].p, fb.elems[p] * n_frames * 2, cudaMemcpyDeviceToHost, s));
  CKS(cudaStreamSynchronize(s));
  return AV1B_OK;
}

int av1b_k_deblock(int device, int width, int height, int bit_depth, int n_frames, const Av1bBlockInfo* blocks,
                   const uint16_t* const in[3], uint16_t* const out[3], const int32_t lf_level[4], int sharpness,
                   int reps, double* ms_per_launch) {
  if (!blocks || !in || !out || !lf_level || n_frames <= 0) { set_error("bad argument"); return AV1B_ERR_INVALID; }
  DeblockLaunch L;
  if (av1b_geom_init(&L.g, width, height, 0, 0)) { set_error("unsupported size"); return AV1B_ERR_INVALID; }
  int rc = select_device(device);
  if (rc) return rc;
  Timer t; CKS(t.init());
  FrameBufs bi, bo;
  if ((rc = upload_planes(L.g, n_frames, in, bi, t.s))) return rc;
  if ((rc = upload_planes(L.g, n_frames, nullptr, bo, t.s))) return rc;
  DevBuf db;
  L.map_elems = (size_t)L.g.w8 * L.g.h8;
  CKS(db.alloc(L.map_elems * n_frames * sizeof(Av1bBlockInfo)));
  CKS(cudaMemcpyAsync(db.p, blocks, L.map_elems * n_frames * sizeof(Av1bBlockInfo), cudaMemcpyHostToDevice, t.s));
  L.bit_depth = bit_depth; L.sharpness = sharpness;
  for (int i = 0; i < 4; i++) L.lf_level[i] = lf_level[i];
  for (int p = 0; p < 3; p++) { L.in[p] = bi.d[p].as<uint16_t>(); L.out[p] = bo.d[p].as<uint16_t>(); L.plane_elems[p] = bi.elems[p]; }
  L.blocks = db.as<Av1bBlockInfo>();
  if ((rc = timed(t, reps, ms_per_launch, [&]() { return launch_deblock(L, n_frames, t.s); }))) return rc;
  return download_planes(n_frames, out, bo, t.s);
}

int av1b_k_cdef(int device, int width, int height, int bit_depth, int n_frames, const Av1bBlockInfo* blocks,
                const Av1bFrameParams* fp, const uint16_t* const in[3], const uint16_t* const src[3],
                const uint8_t* forced_idx, uint16_t* const out[3], uint8_t* cdef_idx_out, int reps,
                double* ms_per_launch) {
  if (!blocks || !fp || !in || !out || n_frames <= 0 || (!src && !forced_idx)) { set_error("bad argument"); return AV1B_ERR_INVALID; }
  CdefLaunch L;
  if (av1b_geom_init(&L.g, width, height, 0, 0)) { set_error("unsupported size"); return AV1B_ERR_INVALID; }
  int rc = select_device(device);
  if (rc) return rc;
  Timer t; CKS(t.init());
  FrameBufs bi, bs, bo;
  if ((rc = upload_planes(L.g, n_frames, in, bi, t.s))) return rc;
  if (src && (rc = upload_planes(L.g, n_frames, src, bs, t.s))) return rc;
  if ((rc = upload_planes(L.g, n_frames, nullptr, bo, t.s))) return rc;
  DevBuf db, dix, dfo;
  L.map_elems = (size_t)L.g.w8 * L.g.h8;
  const size_t nsb = (size_t)L.g.sb_rows * L.g.sb_cols * n_frames;
  CKS(db.alloc(L.map_elems * n_frames * sizeof(Av1bBlockInfo)));
  CKS(cudaMemcpyAsync(db.p, blocks, L.map_elems * n_frames * sizeof(Av1bBlockInfo), cudaMemcpyHostToDevice, t.s));
  CKS(dix.alloc(nsb));
  L.forced_idx = nullptr;
  if (forced_idx) {
    CKS(dfo.alloc(nsb));
    CKS(cudaMemcpyAsync(dfo.p, forced_idx, nsb, cudaMemcpyHostToDevice, t.s));
    L.forced_idx = dfo.as<uint8_t>();
  }
  L.bit_depth = bit_depth; L.cdef_damping = fp->cdef_damping; L.cdef_bits = fp->cdef_bits;
  for (int i = 0; i < 8; i++) { L.y_strength[i] = fp->cdef_y_strength[i]; L.uv_strength[i] = fp->cdef_uv_strength[i]; }
  for (int p = 0; p < 3; p++) {
    L.in[p] = bi.d[p].as<uint16_t>(); L.src[p] = src ? bs.d[p].as<uint16_t>() : bi.d[p].as<uint16_t>();
    L.out[p] = bo.d[p].as<uint16_t>(); L.plane_elems[p] = bi.elems[p];
  }
  L.blocks = db.as<Av1bBlockInfo>();
  L.cdef_idx = dix.as<uint8_t>();
  if ((rc = timed(t, reps, ms_per_launch, [&]() { return launch_cdef(L, n_frames, t.s); }))) return rc;
  if (cdef_idx_out) CKS(cudaMemcpyAsync(cdef_idx_out, dix.p, nsb, cudaMemcpyDeviceToHost, t.s));
  return download_planes(n_frames, out, bo, t.s);
}

int av1b_k_lr(int device, int width, int height, int bit_depth, int n_frames, const Av1bFrameParams* fp,
              const uint16_t* const cdef[3], const uint16_t* const deb[3], const Av1bLrUnit* const units[3],
              uint16_t* const out[3], int reps, double* ms_per_launch) {
  if (!fp || !cdef || !deb || !units || !out || n_frames <= 0) { set_error("bad argument"); return AV1B_ERR_INVALID; }
  LrLaunch L;
  if (av1b_geom_init(&L.g, width, height, 0, 0)) { set_error("unsupported size"); return AV1B_ERR_INVALID; }
  int rc = select_device(device);
  if (rc) return rc;
  Timer t; CKS(t.init());
  FrameBufs bc, bd, bo;
  if ((rc = upload_planes(L.g, n_frames, cdef, bc, t.s))) return rc;
  if ((rc = upload_planes(L.g, n_frames, deb, bd, t.s))) return rc;
  if ((rc = upload_planes(L.g, n_frames, nullptr, bo, t.s))) return rc;
  DevBuf du[3];
  L.bit_depth = bit_depth;
  for (int p = 0; p < 3; p++) {
    const int ss = p > 0;
    int us = 64 << fp->lr_unit_shift;
    if (ss) us >>= fp->lr_uv_shift;
    const int ph = (height + ss) >> ss, pw = (width + ss) >> ss;
    L.lr_type[p] = fp->lr_type[p];
    L.unit_size[p] = us;
    L.unit_rows[p] = (ph + (us >> 1)) / us > 0 ? (ph + (us >> 1)) / us : 1;
    L.unit_cols[p] = (pw + (us >> 1)) / us > 0 ? (pw + (us >> 1)) / us : 1;
    L.units[p] = nullptr;
    if (fp->lr_type[p] != AV1B_RESTORE_NONE && units[p]) {
      const size_t n = (size_t)L.unit_rows[p] * L.unit_cols[p] * n_frames * sizeof(Av1bLrUnit);
      CKS(du[p].alloc(n));
      CKS(cudaMemcpyAsync(du[p].p, units[p], n, cudaMemcpyHostToDevice, t.s));
      L.units[p] = du[p].as<Av1bLrUnit>();
    }
    L.cdef[p] = bc.d[p].as<uint16_t>(); L.deb[p] = bd.d[p].as<uint16_t>(); L.out[p] = bo.d[p].as<uint16_t>();
    L.plane_elems[p] = bc.elems[p];
  }
  if ((rc = timed(t, reps, ms_per_launch, [&]() { return launch_lr(L, n_frames, t.s); }))) return rc;
  return download_planes(n_frames, out, bo, t.s);
}

int av1b_k_lr_search(int device, int width, int height, int bit_depth, const Av1bLrUnit* cand,
                     const uint16_t* const cdef[3], const uint16_t* const deb[3], const uint16_t* src_y, int64_t bias,
                     Av1bLrUnit* units_out, uint64_t* sse_out, int reps, double* ms_per_launch) {
  if (!cand || !cdef || !deb || !src_y || !units_out) { set_error("bad argument"); return AV1B_ERR_INVALID; }
  LrLaunch L;
  if (av1b_geom_init(&L.g, width, height, 0, 0)) { set_error("unsupported size"); return AV1B_ERR_INVALID; }
  int rc = select_device(device);
  if (rc) return rc;
  Timer t; CKS(t.init());
  FrameBufs bc, bd, bs;
  const uint16_t* const srcs[3] = {src_y, cdef[1], cdef[2]};
  if ((rc = upload_planes(L.g, 1, cdef, bc, t.s))) return rc;
  if ((rc = upload_planes(L.g, 1, deb, bd, t.s))) return rc;
  if ((rc = upload_planes(L.g, 1, srcs, bs, t.s))) return rc;
  L.bit_depth = bit_depth;
  for (int p = 0; p < 3; p++) {
    const int ss = p > 0, us = 64 >> ss;
    const int ph = (height + ss) >> ss, pw = (width + ss) >> ss;
    L.lr_type[p] = p ? AV1B_RESTORE_NONE : AV1B_RESTORE_SWITCHABLE;
    L.unit_size[p] = us;
    L.unit_rows[p] = std::max((ph + (us >> 1)) / us, 1);
    L.unit_cols[p] = std::max((pw + (us >> 1)) / us, 1);
    L.units[p] = nullptr;
    L.cdef[p] = bc.d[p].as<uint16_t>(); L.deb[p] = bd.d[p].as<uint16_t>(); L.out[p] = nullptr;
    L.plane_elems[p] = bc.elems[p];
  }
  const size_t n = (size_t)L.unit_rows[0] * L.unit_cols[0];
  DevBuf dsse, du;
  CKS(dsse.alloc(3 * n * sizeof(unsigned long long)));
  CKS(du.alloc(n * sizeof(Av1bLrUnit)));
  L.src_y = bs.d[0].as<uint16_t>(); L.cand = *cand; L.sse = dsse.as<unsigned long long>();
  if ((rc = timed(t, reps, ms_per_launch, [&]() { return launch_lr_search(L, 1, bias, du.as<Av1bLrUnit>(), t.s); }))) return rc;
  CKS(cudaMemcpyAsync(units_out, du.p, n * sizeof(Av1bLrUnit), cudaMemcpyDeviceToHost, t.s));
  if (sse_out) CKS(cudaMemcpyAsync(sse_out, dsse.p, 3 * n * sizeof(uint64_t), cudaMemcpyDeviceToHost, t.s));
  CKS(cudaStreamSynchronize(t.s));
  return AV1B_OK;
}

int av1b_k_pyramid(int device, int width, int height, int n_frames, const uint16_t* l0, uint16_t* l1, uint16_t* l2,
                   int reps, double* ms_per_launch) {
  if (!l0 || !l1 || !l2 || n_frames <= 0) { set_error("bad argument"); return AV1B_ERR_INVALID; }
  Av1bGeom g;
  if (av1b_geom_init(&g, width, height, 0, 0)) { set_error("unsupported size"); return AV1B_ERR_INVALID; }
  int rc = select_device(device);
  if (rc) return rc;
  Timer t; CKS(t.init());
  const size_t e0 = (size_t)g.stride[0] * g.rows[0];
  DevBuf d0, d1, d2;
  CKS(d0.alloc(e0 * n_frames * 2)); CKS(d1.alloc(e0 / 4 * n_frames * 2)); CKS(d2.alloc(e0 / 16 * n_frames * 2));
  CKS(cudaMemcpyAsync(d0.p, l0, e0 * n_frames * 2, cudaMemcpyHostToDevice, t.s));
  if ((rc = timed(t, reps, ms_per_launch, [&]() {
         return launch_pyramid(d0.as<uint16_t>(), d1.as<uint16_t>(), d2.as<uint16_t>(), g.stride[0], g.rows[0], e0, n_frames, t.s);
       }))) return rc;
  CKS(cudaMemcpyAsync(l1, d1.p, e0 / 4 * n_frames * 2, cudaMemcpyDeviceToHost, t.s));
  CKS(cudaMemcpyAsync(l2, d2.p, e0 / 16 * n_frames * 2, cudaMemcpyDeviceToHost, t.s));
  CKS(cudaStreamSynchronize(t.s));
  return AV1B_OK;
}

int av1b_k_hme(int device, int width, int height, int bit_depth, int n_frames, const uint16_t* cur_l0, const uint16_t* ref_l0,
               int lambda, int16_t* mv_out, int reps, double* ms_per_launch) {
  return av1b_k_hme_sbrd(device, width, height, bit_depth, n_frames, cur_l0, ref_l0, lambda, 0, 0, 0, mv_out, reps, ms_per_launch);
}

int av1b_k_hme_sbrd(int device, int width, int height, int bit_depth, int n_frames, const uint16_t* cur_l0, const uint16_t* ref_l0,
                    int lambda, int lam_s, int lam_r, int passes, int16_t* mv_out, int reps, double* ms_per_launch) {
  if (n_frames > kMaxSearches) { set_error("at most %d frames per call", kMaxSearches); return AV1B_ERR_INVALID; }
  if (!cur_l0 || !ref_l0 || !mv_out || n_frames <= 0 || (bit_depth != 8 && bit_depth != 10)) { set_error("bad argument"); return AV1B_ERR_INVALID; }
  Av1bGeom g;
  if (av1b_geom_init(&g, width, height, 0, 0)) { set_error("unsupported size"); return AV1B_ERR_INVALID; }
  int rc = select_device(device);
  if (rc) return rc;
  Timer t; CKS(t.init());
  const size_t e0 = (size_t)g.stride[0] * g.rows[0];
  DevBuf c[3], r[3], m2, mo;
  for (int l = 0; l < 3; l++) { CKS(c[l].alloc((e0 >> (2 * l)) * n_frames * 2)); CKS(r[l].alloc((e0 >> (2 * l)) * n_frames * 2)); }
  const int n2 = ((width + 31) / 32) * ((height + 31) / 32);
  CKS(m2.alloc((size_t)n2 * n_frames * 4));
  CKS(mo.alloc((size_t)g.w8 * g.h8 * n_frames * 4));
  CKS(cudaMemcpyAsync(c[0].p, cur_l0, e0 * n_frames * 2, cudaMemcpyHostToDevice, t.s));
  CKS(cudaMemcpyAsync(r[0].p, ref_l0, e0 * n_frames * 2, cudaMemcpyHostToDevice, t.s));
  CKS(launch_pyramid(c[0].as<uint16_t>(), c[1].as<uint16_t>(), c[2].as<uint16_t>(), g.stride[0], g.rows[0], e0, n_frames, t.s));
  CKS(launch_pyramid(r[0].as<uint16_t>(), r[1].as<uint16_t>(), r[2].as<uint16_t>(), g.stride[0], g.rows[0], e0, n_frames, t.s));
  HmeLaunch L;
  L.width = width; L.height = height; L.stride0 = g.stride[0]; L.elems0 = e0; L.lambda = lambda; L.shift2 = bit_depth - 8;
  for (int l = 0; l < 3; l++) { L.cur[l] = c[l].as<uint16_t>(); L.ref[l] = r[l].as<uint16_t>(); }
  L.mv2 = m2.as<int16_t>(); L.mv_out = mo.as<int16_t>();
  for (int f = 0; f < n_frames; f++) { L.cur_slot[f] = (uint8_t)f; L.ref_slot[f] = (uint8_t)f; }
  DevBuf tmp, hist;
  const size_t n1 = (size_t)((width + 15) / 16) * ((height + 15) / 16);
  CKS(tmp.alloc(n1 * n_frames * 4)); CKS(hist.alloc((size_t)n_frames * 2049 * 4));
  L.lam_s = lam_s; L.lam_r = lam_r; L.sbrd_passes = passes; L.mv_tmp = tmp.as<int16_t>(); L.hist = hist.as<uint32_t>();
  if ((rc = timed(t, reps, ms_per_launch, [&]() {
         cudaError_t e = launch_hme(L, n_frames, t.s);
         return e != cudaSuccess ? e : launch_hme_sbrd(L, n_frames, t.s);
       }))) return rc;
  CKS(cudaMemcpyAsync(mv_out, mo.p, (size_t)g.w8 * g.h8 * n_frames * 4, cudaMemcpyDeviceToHost, t.s));
  CKS(cudaStreamSynchronize(t.s));
  return AV1B_OK;
}

int av1b_k_inter_encode(int device, int width, int height, int bit_depth, int base_q_idx, const uint8_t* part_map,
                        const int16_t* mvs, const uint16_t* const src[3], const uint16_t* const ref[3],
                        uint16_t* const rec[3], int16_t* const coef[3], Av1bBlockInfo* blocks, int tb_zero_thr,
                        int merge_skip, int reps, double* ms_per_launch) {
  if (!part_map || !mvs || !src || !ref || !rec || !coef || !blocks || base_q_idx < 1 || base_q_idx > 255) {
    set_error("bad argument"); return AV1B_ERR_INVALID;
  }
  InterLaunch L;
  if (av1b_geom_init(&L.g, width, height, 0, 0)) { set_error("unsupported size"); return AV1B_ERR_INVALID; }
  int rc = select_device(device);
  if (rc) return rc;
  Timer t; CKS(t.init());
  FrameBufs bs, br, bo;
  if ((rc = upload_planes(L.g, 1, src, bs, t.s))) return rc;
  if ((rc = upload_planes(L.g, 1, ref, br, t.s))) return rc;
  if ((rc = upload_planes(L.g, 1, nullptr, bo, t.s))) return rc;
  const size_t map = (size_t)L.g.w8 * L.g.h8;
  DevBuf dc[3], dpm, dmv, dbl;
  for (int p = 0; p < 3; p++) { CKS(dc[p].alloc(bs.elems[p] * 2)); CKS(cudaMemsetAsync(dc[p].p, 0, bs.elems[p] * 2, t.s)); }
  CKS(dpm.alloc(map)); CKS(dmv.alloc(map * 4)); CKS(dbl.alloc(map * sizeof(Av1bBlockInfo)));
  CKS(cudaMemcpyAsync(dpm.p, part_map, map, cudaMemcpyHostToDevice, t.s));
  CKS(cudaMemcpyAsync(dmv.p, mvs, map * 4, cudaMemcpyHostToDevice, t.s));
  CKS(cudaMemsetAsync(dbl.p, 0, map * sizeof(Av1bBlockInfo), t.s));
  L.bit_depth = bit_depth; L.base_q_idx = base_q_idx; L.quant_rnd = 48;
  L.dc_q = bit_depth == 8 ? av1t_dc_q_8[base_q_idx] : av1t_dc_q_10[base_q_idx];
  L.ac_q = bit_depth == 8 ? av1t_ac_q_8[base_q_idx] : av1t_ac_q_10[base_q_idx];
  for (int p = 0; p < 3; p++) {
    L.src[p] = bs.d[p].as<uint16_t>(); L.ref[p] = br.d[p].as<uint16_t>(); L.rec[p] = bo.d[p].as<uint16_t>();
    L.coef[p] = dc[p].as<int16_t>();
  }
  L.blocks = dbl.as<Av1bBlockInfo>(); L.part_map = dpm.as<uint8_t>(); L.mvs = dmv.as<int16_t>(); L.tb_zero_thr = tb_zero_thr; L.pack_levels = 0;
  L.n_frames = 1; L.map_elems = map;
  for (int p = 0; p < 3; p++) { L.plane_elems[p] = bs.elems[p]; L.digest[p] = nullptr; }
  if ((rc = timed(t, reps, ms_per_launch, [&]() { return launch_inter_encode(L, t.s); }))) return rc;
  if (merge_skip) CKS(launch_merge_skip(L.g, L.blocks, map, 1, t.s));
  for (int p = 0; p < 3; p++) CKS(cudaMemcpyAsync(coef[p], dc[p].p, bs.elems[p] * 2, cudaMemcpyDeviceToHost, t.s));
  CKS(cudaMemcpyAsync(blocks, dbl.p, map * sizeof(Av1bBlockInfo), cudaMemcpyDeviceToHost, t.s));
  return download_planes(1, rec, bo, t.s);
}

int av1b_k_mctf(int device, int width, int height, int bit_depth, const uint16_t* const cur[3], int n_nb,
                const uint16_t* const* nb_planes, const int16_t* const* mvs, int thr_b, int thr_p, uint16_t* const out[3],
                int reps, double* ms_per_launch) {
  if (!cur || !out || n_nb < 0 || n_nb > kMaxNb || (n_nb && (!nb_planes || !mvs)) || thr_b < 1 || thr_p < 1) { set_error("bad argument"); return AV1B_ERR_INVALID; }
  MctfLaunch L;
  if (av1b_geom_init(&L.g, width, height, 0, 0)) { set_error("unsupported size"); return AV1B_ERR_INVALID; }
  int rc = select_device(device);
  if (rc) return rc;
  Timer t; CKS(t.init());
  FrameBufs bc, bo, bn[kMaxNb];
  DevBuf dmv[kMaxNb];
  if ((rc = upload_planes(L.g, 1, cur, bc, t.s))) return rc;
  if ((rc = upload_planes(L.g, 1, nullptr, bo, t.s))) return rc;
  const size_t map = (size_t)L.g.w8 * L.g.h8;
  for (int k = 0; k < n_nb; k++) {
    if ((rc = upload_planes(L.g, 1, nb_planes + 3 * k, bn[k], t.s))) return rc;
    CKS(dmv[k].alloc(map * 4));
    CKS(cudaMemcpyAsync(dmv[k].p, mvs[k], map * 4, cudaMemcpyHostToDevice, t.s));
  }
  L.bit_depth = bit_depth; L.n_nb = n_nb; L.thr_b = thr_b; L.thr_p = thr_p;
  for (int p = 0; p < 3; p++) { L.cur[p] = bc.d[p].as<uint16_t>(); L.out[p] = bo.d[p].as<uint16_t>(); }
  for (int k = 0; k < kMaxNb; k++) {
    for (int p = 0; p < 3; p++) L.nb[k][p] = k < n_nb ? bn[k].d[p].as<uint16_t>() : nullptr;
    L.mvs[k] = k < n_nb ? dmv[k].as<int16_t>() : nullptr;
  }
  if ((rc = timed(t, reps, ms_per_launch, [&]() { return launch_mctf(L, t.s); }))) return rc;
  return download_planes(1, out, bo, t.s);
}

int av1b_k_noise_estimate(int device, int width, int height, const uint16_t* src_y, int32_t* noise_out) {
  if (!src_y || !noise_out) { set_error("bad argument"); return AV1B_ERR_INVALID; }
  Av1bGeom g;
  if (av1b_geom_init(&g, width, height, 0, 0)) { set_error("unsupported size"); return AV1B_ERR_INVALID; }
  int rc = select_device(device);
  if (rc) return rc;
  Timer t; CKS(t.init());
  const size_t e0 = (size_t)g.stride[0] * g.rows[0];
  DevBuf d0, dh;
  CKS(d0.alloc(e0 * 2)); CKS(dh.alloc(4096 * sizeof(uint32_t)));
  CKS(cudaMemcpyAsync(d0.p, src_y, e0 * 2, cudaMemcpyHostToDevice, t.s));
  CKS(launch_noise_hist(g, d0.as<uint16_t>(), dh.as<uint32_t>(), t.s));
  std::vector<uint32_t> hist(4096);
  CKS(cudaMemcpyAsync(hist.data(), dh.p, 4096 * sizeof(uint32_t), cudaMemcpyDeviceToHost, t.s));
  CKS(cudaStreamSynchronize(t.s));
  *noise_out = av1b_noise_from_hist(hist.data());
  return AV1B_OK;
}

int av1b_k_partition_smooth(int device, int width, int height, const uint16_t* src_y, int thr, uint8_t* map_out) {
  if (!src_y || !map_out) { set_error("bad argument"); return AV1B_ERR_INVALID; }
  Av1bGeom g;
  if (av1b_geom_init(&g, width, height, 0, 0)) { set_error("unsupported size"); return AV1B_ERR_INVALID; }
  int rc = select_device(device);
  if (rc) return rc;
  Timer t; CKS(t.init());
  const size_t e0 = (size_t)g.stride[0] * g.rows[0], map = (size_t)g.w8 * g.h8;
  DevBuf d0, dm;
  CKS(d0.alloc(e0 * 2)); CKS(dm.alloc(map));
  CKS(cudaMemcpyAsync(d0.p, src_y, e0 * 2, cudaMemcpyHostToDevice, t.s));
  CKS(launch_partition_smooth(g, d0.as<uint16_t>(), e0, map, thr, dm.as<uint8_t>(), 1, t.s));
  CKS(cudaMemcpyAsync(map_out, dm.p, map, cudaMemcpyDeviceToHost, t.s));
  CKS(cudaStreamSynchronize(t.s));
  return AV1B_OK;
}

}  // extern "C"
