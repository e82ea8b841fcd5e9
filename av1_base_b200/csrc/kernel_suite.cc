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

}  // extern "C"
