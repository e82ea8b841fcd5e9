// C-ABI encoder object: owns the device buffers and streams of ONE GPU and drives, per batch of
// frames, source upload -> device encode kernels -> symbol-stream download -> host entropy coding.
// Boundary replaced: /root/reference/crates/daemon/src/encode/av1an.rs:126-139 (run_av1an).
// There is no CPU fallback: without a CUDA device av1b_encoder_create fails with AV1B_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <chrono>
#include <vector>
#include "../../include/av1b200.h"
#include "av1_tables.h"
#include "bitstream.h"
#include "capi_internal.h"
#include "kernels.cuh"

using namespace av1b;

#define CK(call)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) {                                                          \
      set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));   \
      return AV1B_ERR_CUDA;                                                           \
    }                                                                                 \
  } while (0)

struct KeptFrame {
  std::vector<uint16_t> rec[3];
  std::vector<int16_t> coef[3];
  std::vector<Av1bBlockInfo> blocks;
};

struct av1b_encoder {
  av1b_config cfg;
  Av1bSeqParams seq;
  Av1bGeom g;
  int batch = 0;
  int host_threads = 1;
  int base_q_idx = 0;
  int blk_log2 = 4;
  bool keep = false;
  cudaStream_t stream = nullptr;
  size_t plane_elems[3] = {0, 0, 0};
  size_t map_elems = 0;
  uint16_t* d_src[3] = {nullptr, nullptr, nullptr};
  uint16_t* d_rec[3] = {nullptr, nullptr, nullptr};
  int16_t* d_coef[3] = {nullptr, nullptr, nullptr};
  Av1bBlockInfo* d_blocks = nullptr;
  uint8_t* d_map = nullptr;
  uint16_t* h_src[3] = {nullptr, nullptr, nullptr};   // pinned
  uint16_t* h_rec[3] = {nullptr, nullptr, nullptr};
  int16_t* h_coef[3] = {nullptr, nullptr, nullptr};
  Av1bBlockInfo* h_blocks = nullptr;
  std::vector<KeptFrame> kept;
  // statistics of the last chunk (bench.py reads them through av1b_get_stats)
  double t_h2d_ms = 0, t_kernel_ms = 0, t_d2h_ms = 0, t_pack_ms = 0;
  int64_t kernel_launches = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

extern "C" {

int av1b_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void av1b_config_default(av1b_config* c) {
  memset(c, 0, sizeof(*c));
  c->bit_depth = 10; c->fps_num = 30; c->fps_den = 1;
  c->crf = 30; c->preset = 6; c->keyint = 240; c->lookahead = 0;
  c->tile_cols_log2 = -1; c->tile_rows_log2 = -1;
}

static void free_all(av1b_encoder* e) {
  for (int p = 0; p < 3; p++) {
    cudaFree(e->d_src[p]); cudaFree(e->d_rec[p]); cudaFree(e->d_coef[p]);
    cudaFreeHost(e->h_src[p]); cudaFreeHost(e->h_rec[p]); cudaFreeHost(e->h_coef[p]);
  }
  cudaFree(e->d_blocks); cudaFree(e->d_map); cudaFreeHost(e->h_blocks);
  for (auto& ev : e->ev) if (ev) cudaEventDestroy(ev);
  if (e->stream) cudaStreamDestroy(e->stream);
}

int av1b_encoder_create(const av1b_config* cfg, av1b_encoder** out) {
  if (!cfg || !out) { set_error("null argument"); return AV1B_ERR_INVALID; }
  *out = nullptr;
  if (cfg->bit_depth != 8 && cfg->bit_depth != 10) { set_error("bit_depth must be 8 or 10"); return AV1B_ERR_INVALID; }
  if (cfg->crf < 0 || cfg->crf > 63) { set_error("crf out of range 0..63"); return AV1B_ERR_INVALID; }
  int ndev = av1b_device_count();
  if (ndev <= 0) { set_error("no CUDA device visible (av1b200 has no CPU fallback)"); return AV1B_ERR_NO_DEVICE; }
  if (cfg->device_id < 0 || cfg->device_id >= ndev) { set_error("device_id %d out of range", cfg->device_id); return AV1B_ERR_INVALID; }
  av1b_encoder* e = new av1b_encoder();
  e->cfg = *cfg;
  // tiles: auto = about 4x4 superblocks per tile (tiles x frames-in-flight CTAs fill the 148 SMs)
  Av1bGeom probe;
  if (av1b_geom_init(&probe, cfg->width, cfg->height, 0, 0)) {
    set_error("unsupported frame size %dx%d (multiples of 8, 16..8192 x 16..4352)", cfg->width, cfg->height);
    delete e; return AV1B_ERR_INVALID;
  }
  int tcl = cfg->tile_cols_log2, trl = cfg->tile_rows_log2;
  if (tcl < 0) tcl = av1b_tile_log2(4, probe.sb_cols);
  if (trl < 0) trl = av1b_tile_log2(4, probe.sb_rows);
  av1b_geom_init(&e->g, cfg->width, cfg->height, tcl, trl);
  e->seq.width = cfg->width; e->seq.height = cfg->height; e->seq.bit_depth = cfg->bit_depth;
  e->seq.enable_cdef = 0; e->seq.enable_restoration = 0;
  e->seq.fps_num = cfg->fps_num; e->seq.fps_den = cfg->fps_den; e->seq.color_hdr = cfg->hdr;
  e->base_q_idx = av1t_quantizer_to_qindex[cfg->crf];
  if (e->base_q_idx < 1) e->base_q_idx = 1;   // lossless (qindex 0) is not supported
  e->blk_log2 = cfg->reserved[1] ? cfg->reserved[1] : 4;
  e->keep = cfg->reserved[0] != 0;
  e->batch = cfg->frames_in_flight > 0 ? cfg->frames_in_flight : 8;
  e->host_threads = cfg->host_threads > 0 ? cfg->host_threads : 8;
  if (cudaSetDevice(cfg->device_id) != cudaSuccess) { set_error("cudaSetDevice failed"); delete e; return AV1B_ERR_CUDA; }
  cudaError_t err = cudaSuccess;
  auto A = [&](cudaError_t r) { if (err == cudaSuccess && r != cudaSuccess) err = r; };
  A(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  for (auto& ev : e->ev) A(cudaEventCreate(&ev));
  e->map_elems = (size_t)e->g.w8 * e->g.h8;
  for (int p = 0; p < 3; p++) {
    e->plane_elems[p] = (size_t)e->g.stride[p] * e->g.rows[p];
    const size_t n = e->plane_elems[p] * e->batch;
    A(cudaMalloc(&e->d_src[p], n * 2)); A(cudaMalloc(&e->d_rec[p], n * 2)); A(cudaMalloc(&e->d_coef[p], n * 2));
    A(cudaMallocHost(&e->h_src[p], n * 2)); A(cudaMallocHost(&e->h_rec[p], n * 2)); A(cudaMallocHost(&e->h_coef[p], n * 2));
    if (err == cudaSuccess) { A(cudaMemset(e->d_src[p], 0, n * 2)); A(cudaMemset(e->d_coef[p], 0, n * 2)); A(cudaMemset(e->d_rec[p], 0, n * 2)); }
  }
  A(cudaMalloc(&e->d_blocks, e->map_elems * e->batch * sizeof(Av1bBlockInfo)));
  A(cudaMalloc(&e->d_map, e->map_elems * e->batch));
  A(cudaMallocHost(&e->h_blocks, e->map_elems * e->batch * sizeof(Av1bBlockInfo)));
  if (err != cudaSuccess) {
    set_error("device/pinned allocation failed: %s", cudaGetErrorString(err));
    free_all(e); delete e; return AV1B_ERR_NOMEM;
  }
  *out = e;
  return AV1B_OK;
}

void av1b_encoder_destroy(av1b_encoder* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device_id);
  free_all(e);
  delete e;
}

int av1b_encode_chunk(av1b_encoder* e, const av1b_frame_src* frames, uint32_t n_frames, av1b_packet_cb out_cb,
                      av1b_progress_cb prog_cb, void* user) {
  if (!e || !frames || !out_cb) { set_error("null argument"); return AV1B_ERR_INVALID; }
  CK(cudaSetDevice(e->cfg.device_id));
  const Av1bGeom& g = e->g;
  e->kept.clear();
  e->t_h2d_ms = e->t_kernel_ms = e->t_d2h_ms = e->t_pack_ms = 0;
  e->kernel_launches = 0;
  const auto t_start = std::chrono::steady_clock::now();
  const int bd = e->cfg.bit_depth;
  std::vector<uint8_t> tu;
  for (uint32_t f0 = 0; f0 < n_frames; f0 += e->batch) {
    const int nb = (int)std::min<uint32_t>(e->batch, n_frames - f0);
    // stage + upload the source frames into the padded device planes
    CK(cudaEventRecord(e->ev[0], e->stream));
    for (int b = 0; b < nb; b++) {
      const av1b_frame_src& fs = frames[f0 + b];
      for (int p = 0; p < 3; p++) {
        const int w = p ? g.width >> 1 : g.width, h = p ? g.height >> 1 : g.height;
        uint16_t* hs = e->h_src[p] + (size_t)b * e->plane_elems[p];
        for (int y = 0; y < h; y++) memcpy(hs + (size_t)y * g.stride[p], fs.planes[p] + (size_t)y * fs.stride[p], (size_t)w * 2);
        CK(cudaMemcpyAsync(e->d_src[p] + (size_t)b * e->plane_elems[p], hs, (size_t)g.stride[p] * h * 2,
                           cudaMemcpyHostToDevice, e->stream));
      }
    }
    CK(cudaEventRecord(e->ev[1], e->stream));
    IntraLaunch L;
    L.g = g; L.bit_depth = bd; L.base_q_idx = e->base_q_idx; L.quant_rnd = 48;
    L.dc_q = bd == 8 ? av1t_dc_q_8[e->base_q_idx] : av1t_dc_q_10[e->base_q_idx];
    L.ac_q = bd == 8 ? av1t_ac_q_8[e->base_q_idx] : av1t_ac_q_10[e->base_q_idx];
    for (int p = 0; p < 3; p++) { L.src[p] = e->d_src[p]; L.rec[p] = e->d_rec[p]; L.coef[p] = e->d_coef[p]; L.plane_elems[p] = e->plane_elems[p]; }
    L.blocks = e->d_blocks; L.part_map = e->d_map; L.map_elems = e->map_elems;
    CK(launch_partition_fixed(g, e->blk_log2, e->d_map, nb, e->stream));
    CK(launch_intra_encode(L, nb, e->stream));
    e->kernel_launches += 2;
    CK(cudaEventRecord(e->ev[2], e->stream));
    for (int p = 0; p < 3; p++) {
      CK(cudaMemcpyAsync(e->h_coef[p], e->d_coef[p], e->plane_elems[p] * nb * 2, cudaMemcpyDeviceToHost, e->stream));
      if (e->keep) CK(cudaMemcpyAsync(e->h_rec[p], e->d_rec[p], e->plane_elems[p] * nb * 2, cudaMemcpyDeviceToHost, e->stream));
    }
    CK(cudaMemcpyAsync(e->h_blocks, e->d_blocks, e->map_elems * nb * sizeof(Av1bBlockInfo), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaEventRecord(e->ev[3], e->stream));
    CK(cudaStreamSynchronize(e->stream));
    float ms;
    cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]); e->t_h2d_ms += ms;
    cudaEventElapsedTime(&ms, e->ev[1], e->ev[2]); e->t_kernel_ms += ms;
    cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]); e->t_d2h_ms += ms;
    // host entropy coding + packetisation, in display order
    const auto tp0 = std::chrono::steady_clock::now();
    for (int b = 0; b < nb; b++) {
      Av1bFrameParams fp;
      memset(&fp, 0, sizeof(fp));
      fp.frame_type = AV1B_KEY_FRAME;
      fp.base_q_idx = e->base_q_idx;
      fp.disable_cdf_update = 0;
      fp.tile_cols_log2 = g.tile_cols_log2; fp.tile_rows_log2 = g.tile_rows_log2;
      fp.cdef_damping = 3;
      Av1bFrameSyms sy;
      memset(&sy, 0, sizeof(sy));
      sy.blocks = e->h_blocks + (size_t)b * e->map_elems;
      for (int p = 0; p < 3; p++) { sy.coef[p] = e->h_coef[p] + (size_t)b * e->plane_elems[p]; sy.coef_stride[p] = g.stride[p]; }
      tu.clear();
      write_temporal_delimiter(tu);
      if (f0 + b == 0) write_sequence_header(e->seq, tu);
      if (write_frame(e->seq, fp, g, sy, tu, e->host_threads)) { set_error("write_frame failed"); return AV1B_ERR_INTERNAL; }
      if (e->keep) {
        e->kept.emplace_back();
        KeptFrame& k = e->kept.back();
        for (int p = 0; p < 3; p++) {
          k.rec[p].assign(e->h_rec[p] + (size_t)b * e->plane_elems[p], e->h_rec[p] + (size_t)(b + 1) * e->plane_elems[p]);
          k.coef[p].assign(e->h_coef[p] + (size_t)b * e->plane_elems[p], e->h_coef[p] + (size_t)(b + 1) * e->plane_elems[p]);
        }
        k.blocks.assign(sy.blocks, sy.blocks + e->map_elems);
      }
      if (out_cb(user, tu.data(), tu.size(), (int64_t)(f0 + b), 1)) { set_error("packet callback aborted"); return AV1B_ERR_CALLBACK; }
    }
    const auto tp1 = std::chrono::steady_clock::now();
    e->t_pack_ms += std::chrono::duration<double, std::milli>(tp1 - tp0).count();
    if (prog_cb) {
      const double el = std::chrono::duration<double>(tp1 - t_start).count();
      prog_cb(user, (int64_t)(f0 + nb), (int64_t)n_frames, el > 0 ? (f0 + nb) / el : 0.0);
    }
  }
  return AV1B_OK;
}

int av1b_get_recon(av1b_encoder* e, uint32_t frame, uint16_t* const dst[3], const int32_t stride[3]) {
  if (!e || !dst || !stride) { set_error("null argument"); return AV1B_ERR_INVALID; }
  if (!e->keep || frame >= e->kept.size()) { set_error("recon not kept (set config.reserved[0]=1) or bad index"); return AV1B_ERR_INVALID; }
  const Av1bGeom& g = e->g;
  for (int p = 0; p < 3; p++) {
    const int w = p ? g.width >> 1 : g.width, h = p ? g.height >> 1 : g.height;
    for (int y = 0; y < h; y++) memcpy(dst[p] + (size_t)y * stride[p], e->kept[frame].rec[p].data() + (size_t)y * g.stride[p], (size_t)w * 2);
  }
  return AV1B_OK;
}

// Test/diagnostic access to the device-produced symbol streams of a kept frame (padded layouts).
int av1b_get_frame_syms(av1b_encoder* e, uint32_t frame, Av1bBlockInfo* blocks, int16_t* const coef[3]) {
  if (!e || !e->keep || frame >= e->kept.size()) { set_error("symbols not kept or bad index"); return AV1B_ERR_INVALID; }
  if (blocks) memcpy(blocks, e->kept[frame].blocks.data(), e->map_elems * sizeof(Av1bBlockInfo));
  if (coef) for (int p = 0; p < 3; p++) if (coef[p]) memcpy(coef[p], e->kept[frame].coef[p].data(), e->plane_elems[p] * 2);
  return AV1B_OK;
}

int av1b_get_geom(av1b_encoder* e, Av1bGeom* g) {
  if (!e || !g) return AV1B_ERR_INVALID;
  *g = e->g;
  return AV1B_OK;
}

// stats[0..5] = h2d_ms, kernel_ms, d2h_ms, pack_ms, kernel_launches, base_q_idx  (last chunk)
int av1b_get_stats(av1b_encoder* e, double* stats, int n) {
  if (!e || !stats) return AV1B_ERR_INVALID;
  const double v[6] = {e->t_h2d_ms, e->t_kernel_ms, e->t_d2h_ms, e->t_pack_ms, (double)e->kernel_launches, (double)e->base_q_idx};
  for (int i = 0; i < n && i < 6; i++) stats[i] = v[i];
  return AV1B_OK;
}

}  // extern "C"
