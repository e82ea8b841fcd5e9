// C-ABI encoder object: owns the device buffers and streams of ONE GPU and drives, per batch of frames,
// source upload -> device encode kernels -> tokenizer -> range coding (host pool over downloaded token lists, or
// on the device) -> packets.  Three or four buffer slots rotate: batch k+1 uploads while batch k is on the GPU
// and the batches before it are with the host / the device range coder.
// Boundary replaced: /root/reference/crates/daemon/src/encode/av1an.rs:126-139 (run_av1an).
// There is no CPU fallback: without a CUDA device av1b_encoder_create fails with AV1B_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include "../../include/av1b200.h"
#include "av1_tables.h"
#include "av1_qm_tables.h"
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

namespace {

// pyramid ring: kHist pictures before the batch, the carried anchor, then the batch
constexpr int kHist = 2, kSlotCarried = kHist, kSlotFrame0 = kHist + 1;
// anchor period of the one-level hierarchy unless the configuration names one (rate / quality on the 960x544 clip: 6 codes
// 8 to 10 % fewer bytes than 4 at equal PSNR, 8 loses again: the non-reference frames predict over too long a distance)
constexpr int kDefaultGopPeriod = 6;

// Minimal persistent pool: parallel_for blocks the caller (who also works) until all tasks ran.
class ThreadPool {
 public:
  explicit ThreadPool(int n) {
    for (int i = 0; i < n - 1; i++) th_.emplace_back([this] { loop(); });
  }
  ~ThreadPool() {
    { std::lock_guard<std::mutex> l(m_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  void parallel_for(int n, const std::function<void(int)>& fn) {
    if (n <= 0) return;
    {
      std::lock_guard<std::mutex> l(m_);
      fn_ = &fn; n_ = n; next_ = 0; left_ = n; gen_++;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> l(m_);
    done_.wait(l, [this] { return left_ == 0; });
    fn_ = nullptr;
  }
 private:
  void work() {
    for (;;) {
      int i;
      const std::function<void(int)>* f;
      {
        std::lock_guard<std::mutex> l(m_);
        if (!fn_ || next_ >= n_) return;
        i = next_++; f = fn_;
      }
      (*f)(i);
      {
        std::lock_guard<std::mutex> l(m_);
        if (--left_ == 0) done_.notify_all();
      }
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return stop_ || (gen_ != seen && fn_); });
        if (stop_) return;
        seen = gen_;
      }
      work();
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int n_ = 0, next_ = 0, left_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

struct KeptFrame {
  std::vector<uint16_t> rec[3];
  std::vector<int16_t> coef[3];
  std::vector<Av1bBlockInfo> blocks;
  std::vector<uint8_t> cdef_idx;
  std::vector<Av1bLrUnit> lr_units;
  int is_key = 1;
  int kind = 0;
};

// One batch in flight: device sources and symbol streams, pinned staging of the sources and pinned mirrors of
// what the host reads (token lists or coded tile payloads; levels and block info for key frames).
struct Slot {
  uint16_t* d_src[3] = {nullptr, nullptr, nullptr};
  int16_t* d_coef[3] = {nullptr, nullptr, nullptr};   // device side of the symbol streams (downloaded on the copy stream
  Av1bBlockInfo* d_blocks = nullptr;                  //  while the next batch is being encoded)
  uint8_t* d_cdef_idx = nullptr;
  QualityAcc* d_quality = nullptr;                    // per frame: squared error / SSIM of the reconstruction (config.tune[3])
  QualityAcc* h_quality = nullptr;
  Av1bLrUnit* d_lr_units = nullptr;                   // luma restoration units [batch][lr_n] (loop restoration on)
  Av1bLrUnit* h_lr_units = nullptr;
  // token path (inter frames): packed coefficient symbols, tokenizer scratch, token list + superblock offsets
  uint16_t* d_digest[3] = {nullptr, nullptr, nullptr};
  uint8_t* d_mode_cls = nullptr;
  uint32_t* d_blk_count = nullptr;
  uint32_t* d_sb_off = nullptr;
  uint32_t* h_sb_off = nullptr;
  uint32_t* d_tokens = nullptr;
  uint32_t* h_tokens = nullptr;
  size_t tok_cap = 0;                                 // capacity of d_tokens / h_tokens (tokens)
  TokLaunch tl;                                       // the batch's tokenizer launch (re-run after growing the buffers)
  bool has_tokens = false;
  // device range coder: per-tile regions, byte counts / offsets, contiguous payloads
  uint8_t* d_rc_region = nullptr;
  uint32_t* d_rc_len = nullptr;
  uint32_t* h_rc_len = nullptr;
  uint8_t* d_rc_bytes = nullptr;
  uint8_t* h_rc_bytes = nullptr;
  size_t rc_cap = 0;                                  // capacity of d_rc_bytes / h_rc_bytes
  size_t rc_fetched = 0;
  RcLaunch rl;
  cudaStream_t s_rc = nullptr;                        // the slot's range coder: overlaps the next batches' kernels
  cudaEvent_t ev_rc0 = nullptr, ev_rc1 = nullptr;
  size_t tok_fetched = 0;                             // tokens already downloaded with the offsets
  cudaEvent_t ev_tok0 = nullptr, ev_tok1 = nullptr;
  cudaEvent_t ev_src = nullptr;                       // sources of this slot are resident
  uint32_t* d_score = nullptr;                        // scene-change score of every picture against the one before it
  uint32_t* h_score = nullptr;
  cudaEvent_t ev_score = nullptr;
  bool has_score = false;                             // the scores of the pictures now in d_src have been asked for
  bool staged = false;                                // sources came from the host (statistics)
  bool h2d_pending = false;                           // ev_h2d .. ev_src of the last upload have not been read yet
  uint16_t* h_src[3] = {nullptr, nullptr, nullptr};
  uint16_t* h_rec[3] = {nullptr, nullptr, nullptr};
  int16_t* h_coef[3] = {nullptr, nullptr, nullptr};
  Av1bBlockInfo* h_blocks = nullptr;
  uint8_t* h_cdef_idx = nullptr;
  cudaEvent_t ev_h2d = nullptr, ev_k0 = nullptr, ev_me = nullptr, ev_k1 = nullptr, ev_d2h = nullptr, ev_tf0 = nullptr, ev_tf1 = nullptr;
  std::vector<cudaEvent_t> ev_frame;     // 5 per launch group: before encode, after encode, after deblock, after CDEF, after loop restoration
  std::vector<uint8_t> is_key;
  std::vector<uint8_t> kind;             // per frame: 0 key, 1 anchor, 2 non-reference
  std::vector<std::pair<int, int>> groups;   // (first frame, frames) of every launch group of the batch
  int n_frames = 0;
  int64_t first_index = 0;
};

}  // namespace

struct av1b_encoder {
  av1b_config cfg;
  Av1bSeqParams seq;
  Av1bGeom g;                         // key frames: many small tiles (one CTA per tile in the intra kernel)
  Av1bGeom g_inter;                   // inter frames: few large tiles (longer CDF adaptation, less host overhead)
  int batch = 0;
  int base_q_idx = 0;                 // inter frames
  int base_q_idx_key = 0;             // key frames: 3/4 of it (a better reference pays for itself over the GOP)
  int blk_log2 = 4;
  int keyint = 240;
  bool keep = false;
  bool loop_filters = true;
  bool intra_only = false;
  Av1bFrameParams fp_key, fp_inter;   // frame-level parameters (levels / strengths from the quantiser); fp_inter: anchors
  Av1bFrameParams fp_nonref;          // the frames between two anchors (refresh_frame_flags = 0)
  int gop_period = kDefaultGopPeriod; // every gop_period-th frame after a key frame is an anchor (1: plain P chain)
  int base_q_idx_nonref = 0;
  bool me_smooth = true;              // superblock-level rate-distortion regularisation of the vector field after the search
  bool key_var_part = true;           // key frames: 64x64 / 32x32 blocks where the source is smooth
  bool quality_on = false;            // config.tune[3]: PSNR / SSIM of every frame (progress events)
  double q_sse = 0, q_ssim = 0, q_blocks = 0, q_frames = 0, q_psnr_sum = 0;
  int pipe_next = 0, pipe_done = 0;   // batches launched / finished since the encoder was created (slot = index mod n_slots)
  av1b_packet_cb last_cb = nullptr;   // of the last av1b_encode_stream call (packets still in flight belong to it)
  void* last_user = nullptr;
  int q_nominal = 0;                  // quantiser index of the CRF
  bool gop_auto = true;               // structure chosen per chunk from the source's noise level (config.gop_period == 0)
  bool mctf_cfg = true;               // temporal filter allowed by the configuration
  int noise_b = 0;                    // noise estimate of the chunk's first picture (av1b_noise_from_hist)
  bool scene_on = true;               // key frame at a scene change inside a chunk (config.tune[4] = 1: off)
  int64_t scene_cuts = 0;             // key frames the scene scores put inside chunks (statistics)
  long long sc_level = -1;            // running level of change (scene scores), -1: none yet
  int grain_scaling = 0;              // film grain synthesis strength of the chunk (--film-grain > 0 and a filtered structure)
  uint32_t* d_noise_hist = nullptr;
  int src_w = 0, src_h = 0;           // size of the pictures handed in (config.width / height); g.width / g.height = the coded size
  bool padded_src = false;            // the source is not a multiple of 8: stage() pads it by edge replication
  uint8_t* d_qm = nullptr;            // --enable-qm: av1t_qm_sq on the device ([15 levels][luma, chroma][1360])
  uint32_t* h_noise_hist = nullptr;
  cudaEvent_t ev_noise = nullptr;
  int cdf_q[2] = {-1, -1};            // quantisers the device CDF images were made for
  bool mctf_on = true;                // key / anchor source pictures are temporally filtered before they are coded
  int mctf_radius = 2, mctf_key_fwd = 4;
  int hist_count = 0;                 // source pictures of the previous batch that are kept (same GOP): at most kHist
  uint16_t* d_hist_src[3] = {nullptr, nullptr, nullptr};   // [kHist] source pictures before the batch (temporal filter)
  uint16_t* d_flt[3] = {nullptr, nullptr, nullptr};        // [batch] filtered key / anchor sources
  int16_t* d_mvs_tf = nullptr;        // vectors of the temporal filter's searches [kMaxSearches][map_elems][2]
  int16_t* d_mv2_tf = nullptr;
  int64_t mctf_frames = 0;
  double t_mctf_ms = 0;
  uint16_t* d_clip[3] = {nullptr, nullptr, nullptr};       // resident clip (av1b_stage_clip): n_clip source pictures in HBM
  uint32_t n_clip = 0;
  int16_t* d_mv_tmp = nullptr;
  uint32_t* d_hist = nullptr;
  void* d_cdf_init_alt = nullptr;     // default CDF set of the non-reference frames' quantiser class
  cudaStream_t stream = nullptr;      // compute
  cudaStream_t s_in = nullptr;        // host -> device copies (overlap the previous batch's kernels)
  cudaStream_t s_out = nullptr;       // device -> host copies (overlap the next batch's kernels)
  size_t plane_elems[3] = {0, 0, 0};
  size_t map_elems = 0;
  // device working set of one batch (single copy: all device work is ordered on one stream)
  uint16_t* d_rec[3] = {nullptr, nullptr, nullptr};   // reconstruction before the loop filters   [batch]
  uint16_t* d_deb[3] = {nullptr, nullptr, nullptr};   // deblocked                                 [batch]
  uint16_t* d_fin[3] = {nullptr, nullptr, nullptr};   // decoder output; [0] = last frame of the previous batch [batch+1]
  uint8_t* d_map_key = nullptr;
  uint8_t* d_map_inter = nullptr;
  uint16_t* d_pyr[3] = {nullptr, nullptr, nullptr};   // luma pyramid levels 0..2; [0] = last frame of the previous batch
  int16_t* d_mv2 = nullptr;
  int16_t* d_mvs = nullptr;
  bool lr_on = false;                 // loop restoration with per-unit decision (preset <= 5)
  int lr_rows = 0, lr_cols = 0;
  size_t lr_n = 0;
  Av1bLrUnit lr_cand;                 // the Wiener taps / self-guided parameters the units choose from
  unsigned long long* d_lr_sse = nullptr;
  bool token_path = true;             // inter frames: device tokenizer + host range coder over tokens
  int legacy_pack_levels = 0;         // token_path off: 0 raster levels, 1 in-place packed symbols
  uint32_t* d_sb_of_order = nullptr;  // inter-frame tile layout: coding order -> superblock
  uint16_t* d_tile_of_sb = nullptr;
  std::vector<uint32_t> tile_first_k; // [tiles + 1] first coding-order index of each inter-frame tile
  bool rc_on = true;                  // range coding of the token lists on the device (else on the host pool)
  void* d_cdf_init = nullptr;         // default CDF set of the inter frames' quantiser class
  uint32_t* d_tile_first_k = nullptr;
  uint32_t* d_rc_overflow = nullptr;
  size_t rc_guess = 0;
  double t_rc_ms = 0;
  cudaStream_t s_tok = nullptr;       // token list download (issued once the batch's total is known)
  Slot slot[4];                       // batch k+1 uploads / batch k on the GPU / the batches before with the host or the range coder
  int n_slots = 3;                    // 4 with the device range coder: its latency (the largest tile) may span two more batches
  size_t tok_guess = 0;               // tokens of the last batch: that much is downloaded before the total is known
  ThreadPool* pool = nullptr;
  int host_threads = 1;
  int64_t chunk_pos = 0;              // frames since the last key frame
  std::vector<KeptFrame> kept;
  // statistics of the last chunk / resident run
  double t_h2d_ms = 0, t_kernel_ms = 0, t_intra_ms = 0, t_inter_ms = 0, t_me_ms = 0, t_d2h_ms = 0, t_pack_ms = 0,
         t_deblock_ms = 0, t_cdef_ms = 0, t_tok_ms = 0, t_lr_ms = 0;
  int64_t n_tokens = 0, d2h_bytes = 0;
  int64_t kernel_launches = 0, intra_launches = 0, inter_launches = 0, frames_done = 0, bytes_out = 0, key_frames = 0, staged_direct = 0;
};

static void free_all(av1b_encoder* e) {
  for (auto& s : e->slot) {
    for (int p = 0; p < 3; p++) {
      cudaFree(s.d_src[p]); cudaFree(s.d_coef[p]);
      cudaFreeHost(s.h_src[p]); cudaFreeHost(s.h_rec[p]); cudaFreeHost(s.h_coef[p]);
    }
    cudaFreeHost(s.h_blocks); cudaFreeHost(s.h_cdef_idx);
    cudaFree(s.d_blocks); cudaFree(s.d_cdef_idx);
    cudaFree(s.d_lr_units); cudaFreeHost(s.h_lr_units);
    cudaFree(s.d_quality); cudaFreeHost(s.h_quality);
    cudaFree(s.d_score); cudaFreeHost(s.h_score); if (s.ev_score) cudaEventDestroy(s.ev_score);
    for (int p = 0; p < 3; p++) cudaFree(s.d_digest[p]);
    if (s.s_rc) cudaStreamDestroy(s.s_rc);
    cudaFree(s.d_rc_region); cudaFree(s.d_rc_len); cudaFree(s.d_rc_bytes); cudaFreeHost(s.h_rc_len); cudaFreeHost(s.h_rc_bytes);
    cudaFree(s.d_mode_cls); cudaFree(s.d_blk_count); cudaFree(s.d_sb_off); cudaFree(s.d_tokens);
    cudaFreeHost(s.h_sb_off); cudaFreeHost(s.h_tokens);
    for (cudaEvent_t ev : {s.ev_h2d, s.ev_k0, s.ev_me, s.ev_k1, s.ev_d2h, s.ev_src, s.ev_tok0, s.ev_tok1, s.ev_rc0, s.ev_rc1, s.ev_tf0, s.ev_tf1}) if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : s.ev_frame) if (ev) cudaEventDestroy(ev);
  }
  for (int p = 0; p < 3; p++) {
    cudaFree(e->d_rec[p]); cudaFree(e->d_deb[p]); cudaFree(e->d_fin[p]); cudaFree(e->d_pyr[p]);
  }
  cudaFree(e->d_map_key); cudaFree(e->d_map_inter);
  cudaFree(e->d_mv2); cudaFree(e->d_mvs);
  cudaFree(e->d_sb_of_order); cudaFree(e->d_tile_of_sb); cudaFree(e->d_lr_sse);
  cudaFree(e->d_cdf_init); cudaFree(e->d_cdf_init_alt); cudaFree(e->d_tile_first_k); cudaFree(e->d_rc_overflow);
  cudaFree(e->d_mv_tmp); cudaFree(e->d_hist); cudaFree(e->d_mvs_tf); cudaFree(e->d_mv2_tf);
  cudaFree(e->d_qm);
  cudaFree(e->d_noise_hist); cudaFreeHost(e->h_noise_hist); if (e->ev_noise) cudaEventDestroy(e->ev_noise);
  for (int p = 0; p < 3; p++) { cudaFree(e->d_hist_src[p]); cudaFree(e->d_flt[p]); cudaFree(e->d_clip[p]); }
  if (e->s_tok) cudaStreamDestroy(e->s_tok);
  if (e->stream) cudaStreamDestroy(e->stream);
  if (e->s_in) cudaStreamDestroy(e->s_in);
  if (e->s_out) cudaStreamDestroy(e->s_out);
  delete e->pool;
}

// true when the whole plane lies in page-locked memory the device can read directly (av1b_host_alloc,
// cudaHostAlloc / cudaHostRegister of the caller)
static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// upload time of the slot's last staging, once its events have completed (the slot may have been staged again for a later
// batch before the batch that used it is finished: then the reading waits for the next call)
static void collect_h2d(av1b_encoder* e, Slot& s) {
  if (!s.h2d_pending) return;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, s.ev_h2d, s.ev_src) == cudaSuccess) { e->t_h2d_ms += ms; s.h2d_pending = false; }
  else cudaGetLastError();
}

// upload n frames (host pointers) into a slot on the input copy stream (overlaps the kernels of the
// previous batch).  Page-locked sources are read by the copy engine where they lie; pageable ones are
// first gathered into the slot's pinned staging buffer by the host pool.
static int stage(av1b_encoder* e, Slot& s, const av1b_frame_src* frames, int n) {
  const Av1bGeom& g = e->g;
  collect_h2d(e, s);
  s.h2d_pending = true;
  s.has_score = false;
  bool direct = !e->padded_src;        // a padded source is completed in the staging buffer
  for (int b = 0; b < n && direct; b++)
    for (int p = 0; p < 3 && direct; p++) direct = is_pinned(frames[b].planes[p]);
  if (direct) {
    CK(cudaEventRecord(s.ev_h2d, e->s_in));
    for (int b = 0; b < n; b++)
      for (int p = 0; p < 3; p++) {
        const int w = p ? g.width >> 1 : g.width, h = p ? g.height >> 1 : g.height;
        CK(cudaMemcpy2DAsync(s.d_src[p] + (size_t)b * e->plane_elems[p], (size_t)g.stride[p] * 2, frames[b].planes[p],
                             (size_t)frames[b].stride[p] * 2, (size_t)w * 2, h, cudaMemcpyHostToDevice, e->s_in));
      }
    CK(cudaEventRecord(s.ev_src, e->s_in));
    e->staged_direct += n;
    return AV1B_OK;
  }
  CK(cudaEventSynchronize(s.ev_src));   // the previous upload out of this slot's staging buffer has finished
  for (int p = 0; p < 3; p++)
    if (!s.h_src[p]) {
      const size_t bytes = e->plane_elems[p] * (size_t)e->batch * 2;
      CK(cudaMallocHost(&s.h_src[p], bytes));
      memset(s.h_src[p], 0, bytes);
    }
  const int kSplit = 4;   // row bands per plane
  e->pool->parallel_for(n * 3 * kSplit, [&](int task) {
    const int b = task / (3 * kSplit), p = (task / kSplit) % 3, band = task % kSplit;
    // w x h: the plane as handed in; cw x ch: the coded plane (equal unless the source is padded)
    const int w = p ? (e->src_w + 1) >> 1 : e->src_w, h = p ? (e->src_h + 1) >> 1 : e->src_h;
    const int cw = p ? g.width >> 1 : g.width, ch = p ? g.height >> 1 : g.height;
    const int y0 = h * band / kSplit, y1 = h * (band + 1) / kSplit;
    uint16_t* hs = s.h_src[p] + (size_t)b * e->plane_elems[p];
    const uint16_t* sp = frames[b].planes[p];
    const int sst = frames[b].stride[p];
    for (int y = y0; y < y1; y++) {
      uint16_t* row = hs + (size_t)y * g.stride[p];
      memcpy(row, sp + (size_t)y * sst, (size_t)w * 2);
      for (int x = w; x < cw; x++) row[x] = row[w - 1];
    }
    if (band == kSplit - 1)
      for (int y = h; y < ch; y++) memcpy(hs + (size_t)y * g.stride[p], hs + (size_t)(h - 1) * g.stride[p], (size_t)cw * 2);
  });
  CK(cudaEventRecord(s.ev_h2d, e->s_in));
  for (int p = 0; p < 3; p++) {
    const int h = p ? g.height >> 1 : g.height;
    // one strided copy per plane (n frames): the padded rows below the picture stay zero on both sides
    CK(cudaMemcpy2DAsync(s.d_src[p], e->plane_elems[p] * 2, s.h_src[p], e->plane_elems[p] * 2,
                         (size_t)g.stride[p] * h * 2, n, cudaMemcpyHostToDevice, e->s_in));
  }
  CK(cudaEventRecord(s.ev_src, e->s_in));
  return AV1B_OK;
}

// kind of the frame at position `pos` of its closed GOP: 0 key, 1 anchor (inter frame that becomes the reference),
// 2 non-reference inter frame (predicts from the last anchor, refresh_frame_flags = 0)
static int frame_kind(const av1b_encoder* e, int64_t pos) {
  if (e->intra_only) return 0;
  const int64_t c = pos % e->keyint;
  if (c == 0) return 0;
  return (e->gop_period <= 1 || c % e->gop_period == 0) ? 1 : 2;
}
// --enable-qm 1 --qm-min A --qm-max B (av1an.rs:14 passes 1 / 1 / 15): quantisation matrices at the level SVT-AV1 / libaom map the
// frame's quantiser index to (aom_get_qmlevel: min + qindex * (max + 1 - min) / 256), luma and chroma alike; 15 = flat
static void set_qm_levels(const av1b_encoder* e, Av1bFrameParams* fp) {
  const av1b_config& c = e->cfg;
  fp->using_qmatrix = c.enable_qm ? 1 : 0;
  const int lvl = c.enable_qm ? std::min(15, std::max(0, c.qm_min + (fp->base_q_idx * (c.qm_max + 1 - c.qm_min)) / 256)) : 15;
  fp->qm_level[0] = fp->qm_level[1] = lvl;
}
// the device matrices of plane class pc (0 luma, 1 chroma) for a frame, nullptr = flat
static const uint8_t* qm_ptr(const av1b_encoder* e, const Av1bFrameParams& fp, int pc) {
  if (!fp.using_qmatrix || !e->d_qm || fp.qm_level[pc] >= 15) return nullptr;
  return e->d_qm + ((size_t)fp.qm_level[pc] * 2 + pc) * AV1T_QM_SQ_SIZE;
}
static const Av1bFrameParams& kind_params(const av1b_encoder* e, int kind) {
  return kind == 0 ? e->fp_key : (kind == 1 ? e->fp_inter : e->fp_nonref);
}

// device work + symbol download for the n frames resident in the slot; asynchronous
static int launch(av1b_encoder* e, Slot& s, const Slot& in, int n, int64_t first_index) {
  const Av1bGeom& g = e->g;
  const int bd = e->cfg.bit_depth;
  const size_t nsb = (size_t)g.sb_rows * g.sb_cols;
  s.n_frames = n; s.first_index = first_index;
  s.is_key.assign(n, 1); s.kind.assign(n, 0); s.groups.clear();
  // ring index of every frame's reference picture: 0 = the anchor carried over from the batches before, b + 1 = frame b
  // of this batch (d_fin / d_pyr hold batch + 1 pictures)
  std::vector<int> ref_of(n, 0);
  bool any_inter = false;
  int last_ref = 0;
  // position of every frame in its closed GOP: it restarts at a scene change the scores of this batch show (the scores were
  // computed on the upload stream when the pictures arrived, normally a batch ago: no stall); integer rule, see orc_scene_score
  std::vector<int64_t> gop_pos(n);
  {
    const bool use = in.has_score && e->scene_on && !e->intra_only;
    if (use) CK(cudaEventSynchronize(in.ev_score));
    if (e->chunk_pos == 0) e->sc_level = -1;
    const long long unit = (long long)((g.width - 4 + 7) / 8) * ((g.height - 4 + 7) / 8) << (bd - 8);
    int64_t pos = e->chunk_pos;
    for (int b = 0; b < n; b++) {
      if (use && !(b == 0 && e->chunk_pos == 0)) {
        const long long sc = in.h_score[b];
        const bool cut = pos >= 12 && sc > 10 * unit && (e->sc_level < 0 || sc > 3 * e->sc_level + 2 * unit);
        e->sc_level = e->sc_level < 0 ? sc : (4 * e->sc_level + sc) / 5;
        if (cut) { e->sc_level = -1; pos = 0; e->scene_cuts++; }
      }
      gop_pos[b] = pos++;
    }
  }
  for (int b = 0; b < n; b++) {
    const int kind = frame_kind(e, gop_pos[b]);
    s.kind[b] = (uint8_t)kind; s.is_key[b] = kind == 0;
    ref_of[b] = last_ref;
    if (kind != 2) last_ref = b + 1;
    any_inter |= kind != 0;
  }
  CK(cudaStreamWaitEvent(e->stream, in.ev_src, 0));
  CK(cudaEventRecord(s.ev_k0, e->stream));
  const int acq0 = bd == 8 ? av1t_ac_q_8[e->base_q_idx] : av1t_ac_q_10[e->base_q_idx];
  // ---- open-loop motion estimation for the whole batch (source pictures only) ----
  std::vector<uint8_t> filtered(n, 0);
  if (!e->intra_only) {
    const size_t e0 = e->plane_elems[0];
    if (e->chunk_pos == 0) e->hist_count = 0;   // a new closed GOP: nothing before it may be looked at
    auto lvl = [&](int l, int slot) { return e->d_pyr[l] + (e0 >> (2 * l)) * (size_t)slot; };
    CK(cudaMemcpyAsync(lvl(0, kSlotFrame0), in.d_src[0], e0 * n * 2, cudaMemcpyDeviceToDevice, e->stream));
    CK(launch_pyramid(lvl(0, kSlotFrame0), lvl(1, kSlotFrame0), lvl(2, kSlotFrame0), g.stride[0], g.rows[0], e0, n, e->stream));
    e->kernel_launches += 1;
    HmeLaunch H;
    H.width = g.width; H.height = g.height; H.stride0 = g.stride[0]; H.elems0 = e0;
    for (int l = 0; l < 3; l++) { H.ref[l] = e->d_pyr[l]; H.cur[l] = e->d_pyr[l]; }
    H.shift2 = bd - 8;
    H.lambda = acq0 >> 1;   // vector-deviation cost in SAD units (tuned on the oracle: about half the AC quantiser step)
    H.lam_s = e->me_smooth ? H.lambda : 0; H.lam_r = H.lambda >> 2; H.sbrd_passes = e->cfg.preset <= 3 ? 3 : 2; H.mv_tmp = e->d_mv_tmp; H.hist = e->d_hist;
    if (any_inter) {
      for (int b = 0; b < n; b++) { H.cur_slot[b] = (uint8_t)(kSlotFrame0 + b); H.ref_slot[b] = (uint8_t)(ref_of[b] == 0 ? kSlotCarried : kSlotFrame0 + ref_of[b] - 1); }
      H.mv2 = e->d_mv2; H.mv_out = e->d_mvs;
      CK(launch_hme(H, n, e->stream));
      e->kernel_launches += 2;
      if (H.lam_s > 0) { CK(launch_hme_sbrd(H, n, e->stream)); e->kernel_launches += 2 + 2 * H.sbrd_passes; }
    }
    // ---- temporal filter of the key / anchor sources: searches of the picture against its neighbours in time, then
    //      one filter launch per picture ----
    if (e->mctf_on) {
      CK(cudaEventRecord(s.ev_tf0, e->stream));
      struct Job { int b, first, count; int nb[kMaxNb]; };
      std::vector<Job> jobs;
      int n_pairs = 0;
      for (int b = 0; b < n; b++) {
        if (s.kind[b] == 2) continue;
        const int64_t in_gop = gop_pos[b] % e->keyint;
        int lo = s.kind[b] == 0 ? 0 : -e->mctf_radius, hi = s.kind[b] == 0 ? e->mctf_key_fwd : e->mctf_radius;
        if (e->cfg.lookahead >= 0) hi = std::min(hi, e->cfg.lookahead);
        Job j; j.b = b; j.first = n_pairs; j.count = 0;
        for (int d = lo; d <= hi && j.count < kMaxNb; d++) {
          if (d == 0) continue;
          const int r = b + d;
          if (r >= n || r < -e->hist_count) continue;
          if (in_gop + d < 0 || in_gop + d >= e->keyint) continue;   // stay inside the closed GOP
          if (d > 0 && gop_pos[r] != gop_pos[b] + d) continue;       // (a scene change between the two)
          if (n_pairs >= kMaxSearches) break;
          H.cur_slot[n_pairs] = (uint8_t)(kSlotFrame0 + b);
          H.ref_slot[n_pairs] = (uint8_t)(r >= 0 ? kSlotFrame0 + r : kHist + r);
          j.nb[j.count++] = r; n_pairs++;
        }
        if (j.count) jobs.push_back(j);
      }
      if (n_pairs) {
        H.mv2 = e->d_mv2_tf; H.mv_out = e->d_mvs_tf;
        // (no regularisation sweeps here: they pay for vectors that are coded, the filter gains 0.5 % from them)
        CK(launch_hme(H, n_pairs, e->stream));
        e->kernel_launches += 2;
      }
      for (const Job& j : jobs) {
        const Av1bFrameParams& fpk = kind_params(e, s.kind[j.b]);
        const long long aq = bd == 8 ? av1t_ac_q_8[fpk.base_q_idx] : av1t_ac_q_10[fpk.base_q_idx];
        MctfLaunch M;
        M.g = g; M.bit_depth = bd; M.n_nb = j.count;
        // strength follows the quantiser (what it would quantise away anyway may as well be averaged away) and --film-grain
        M.thr_b = (int)std::max<long long>(1, (aq * aq * (10 + e->cfg.film_grain)) / 2560);
        M.thr_p = 3 * M.thr_b;
        for (int p = 0; p < 3; p++) { M.cur[p] = in.d_src[p] + (size_t)j.b * e->plane_elems[p]; M.out[p] = e->d_flt[p] + (size_t)j.b * e->plane_elems[p]; }
        for (int k = 0; k < kMaxNb; k++) {
          const int r = k < j.count ? j.nb[k] : 0;
          for (int p = 0; p < 3; p++)
            M.nb[k][p] = k >= j.count ? nullptr : (r >= 0 ? in.d_src[p] + (size_t)r * e->plane_elems[p] : e->d_hist_src[p] + (size_t)(kHist + r) * e->plane_elems[p]);
          M.mvs[k] = k < j.count ? e->d_mvs_tf + (size_t)(j.first + k) * e->map_elems * 2 : nullptr;
        }
        CK(launch_mctf(M, e->stream));
        e->kernel_launches += 1; e->mctf_frames += 1;
        filtered[j.b] = 1;
      }
      CK(cudaEventRecord(s.ev_tf1, e->stream));
    }
    // the last key / anchor picture of this batch is the reference the next batch starts from
    if (last_ref > 0)
      for (int l = 0; l < 3; l++)
        CK(cudaMemcpyAsync(lvl(l, kSlotCarried), lvl(l, kSlotFrame0 + last_ref - 1), (e0 >> (2 * l)) * 2, cudaMemcpyDeviceToDevice, e->stream));
    // the last kHist source pictures stay for the next batch's temporal filter (ascending: an old entry moves down first)
    if (e->mctf_on) {
      for (int k = 0; k < kHist; k++) {
        const int rel = n - kHist + k;
        for (int l = 0; l < 3; l++)
          CK(cudaMemcpyAsync(lvl(l, k), lvl(l, rel >= 0 ? kSlotFrame0 + rel : kHist + rel), (e0 >> (2 * l)) * 2, cudaMemcpyDeviceToDevice, e->stream));
        for (int p = 0; p < 3; p++)
          CK(cudaMemcpyAsync(e->d_hist_src[p] + (size_t)k * e->plane_elems[p],
                             rel >= 0 ? in.d_src[p] + (size_t)rel * e->plane_elems[p] : e->d_hist_src[p] + (size_t)(kHist + rel) * e->plane_elems[p],
                             e->plane_elems[p] * 2, cudaMemcpyDeviceToDevice, e->stream));
      }
      e->hist_count = std::min(kHist, e->hist_count + n);
    }
  }
  CK(cudaEventRecord(s.ev_me, e->stream));
  // ---- launch groups: a key frame alone; an anchor alone; the non-reference frames between two anchors together
  //      (same reference, same quantiser, no dependency on each other) ----
  for (int b = 0; b < n;) {
    int cnt = 1;
    if (s.kind[b] == 2) while (b + cnt < n && s.kind[b + cnt] == 2 && ref_of[b + cnt] == ref_of[b]) cnt++;
    s.groups.emplace_back(b, cnt);
    b += cnt;
  }
  for (size_t gi = 0; gi < s.groups.size(); gi++) {
    const int b = s.groups[gi].first, cnt = s.groups[gi].second;
    const int kind = s.kind[b];
    const bool key = kind == 0;
    const Av1bFrameParams& fp = kind_params(e, kind);
    const int qidx = fp.base_q_idx;
    const int dcq = bd == 8 ? av1t_dc_q_8[qidx] : av1t_dc_q_10[qidx];
    const int acq = bd == 8 ? av1t_ac_q_8[qidx] : av1t_ac_q_10[qidx];
    cudaEvent_t* ev = &s.ev_frame[gi * 5];
    CK(cudaEventRecord(ev[0], e->stream));
    uint16_t* rec[3]; uint16_t* deb[3]; uint16_t* fin[3]; const uint16_t* refp[3]; const uint16_t* src[3]; int16_t* coef[3];
    for (int p = 0; p < 3; p++) {
      const size_t off = (size_t)b * e->plane_elems[p];
      rec[p] = e->d_rec[p] + off; deb[p] = e->d_deb[p] ? e->d_deb[p] + off : nullptr;
      refp[p] = e->d_fin[p] + (size_t)ref_of[b] * e->plane_elems[p]; fin[p] = e->d_fin[p] + off + e->plane_elems[p];
      src[p] = (filtered[b] ? e->d_flt[p] : in.d_src[p]) + off; coef[p] = s.d_coef[p] + off;
    }
    Av1bBlockInfo* blocks = s.d_blocks + (size_t)b * e->map_elems;
    if (key) {
      IntraLaunch L;
      L.g = g; L.bit_depth = bd; L.base_q_idx = qidx; L.quant_rnd = 48; L.dc_q = dcq; L.ac_q = acq;
      for (int p = 0; p < 3; p++) { L.src[p] = src[p]; L.rec[p] = e->loop_filters ? rec[p] : fin[p]; L.coef[p] = coef[p]; L.plane_elems[p] = e->plane_elems[p]; }
      L.blocks = blocks; L.part_map = e->d_map_key; L.map_elems = e->map_elems;
      L.qm[0] = qm_ptr(e, fp, 0); L.qm[1] = qm_ptr(e, fp, 1);
      if (e->key_var_part) {
        // 64x64 / 32x32 blocks where the source is smooth at this quantiser, 16x16 elsewhere
        const int thr = std::min(4 * acq, 800 << (bd - 8));
        CK(launch_partition_smooth(g, src[0], e->plane_elems[0], e->map_elems, thr, e->d_map_key, 1, e->stream));
        CK(launch_intra_encode(L, 1, e->stream)); e->kernel_launches += 2;
      } else if (e->blk_log2 == 4) { CK(launch_intra_fast(L, 1, e->stream)); e->kernel_launches += 3; }
      else { CK(launch_intra_encode(L, 1, e->stream)); e->kernel_launches += 1; }
      e->intra_launches += 1; e->key_frames += 1;
    } else {
      InterLaunch L;
      L.g = g; L.bit_depth = bd; L.base_q_idx = qidx; L.quant_rnd = 48; L.dc_q = dcq; L.ac_q = acq;
      L.n_frames = cnt; L.map_elems = e->map_elems;
      for (int p = 0; p < 3; p++) { L.src[p] = src[p]; L.ref[p] = refp[p]; L.rec[p] = e->loop_filters ? rec[p] : fin[p]; L.coef[p] = coef[p]; L.plane_elems[p] = e->plane_elems[p]; }
      L.blocks = blocks; L.part_map = e->d_map_inter; L.mvs = e->d_mvs + (size_t)b * e->map_elems * 2;
      L.pack_levels = e->token_path ? 2 : e->legacy_pack_levels;
      for (int p = 0; p < 3; p++) L.digest[p] = e->token_path ? s.d_digest[p] + (size_t)b * e->plane_elems[p] : nullptr;
      L.tb_zero_thr = e->cfg.reserved[4];   // experiment knob: drop transform blocks with sum|level| <= thr
      L.qm[0] = qm_ptr(e, fp, 0); L.qm[1] = qm_ptr(e, fp, 1);
      CK(launch_inter_encode(L, e->stream));
      CK(launch_merge_skip(g, blocks, e->map_elems, cnt, e->stream));
      e->kernel_launches += 2; e->inter_launches += cnt;   // counts inter FRAMES (a launch codes cnt of them)
    }
    CK(cudaEventRecord(ev[1], e->stream));
    if (e->loop_filters) {
      // frames without CDEF (the non-reference frames): the deblocked picture is the decoder's output
      const bool cdef_on = fp.cdef_bits > 0 || fp.cdef_y_strength[0] > 0 || fp.cdef_uv_strength[0] > 0;
      DeblockLaunch D;
      D.g = g; D.bit_depth = bd; D.sharpness = fp.lf_sharpness;
      for (int i = 0; i < 4; i++) D.lf_level[i] = fp.lf_level[i];
      for (int p = 0; p < 3; p++) { D.in[p] = rec[p]; D.out[p] = (cdef_on || e->lr_on) ? deb[p] : fin[p]; D.plane_elems[p] = e->plane_elems[p]; }
      D.blocks = blocks; D.map_elems = e->map_elems;
      CK(launch_deblock(D, cnt, e->stream));
      CK(cudaEventRecord(ev[2], e->stream));
      if (!cdef_on) CK(cudaMemsetAsync(s.d_cdef_idx + (size_t)b * nsb, 0, nsb * cnt, e->stream));
      CdefLaunch Cd;
      Cd.g = g; Cd.bit_depth = bd; Cd.cdef_damping = fp.cdef_damping; Cd.cdef_bits = fp.cdef_bits;
      for (int i = 0; i < 8; i++) { Cd.y_strength[i] = fp.cdef_y_strength[i]; Cd.uv_strength[i] = fp.cdef_uv_strength[i]; }
      // with loop restoration the CDEF output goes to the (now free) pre-filter buffer and restoration writes the picture
      for (int p = 0; p < 3; p++) { Cd.in[p] = deb[p]; Cd.src[p] = src[p]; Cd.out[p] = e->lr_on ? rec[p] : fin[p]; Cd.plane_elems[p] = e->plane_elems[p]; }
      Cd.blocks = blocks; Cd.map_elems = e->map_elems; Cd.cdef_idx = s.d_cdef_idx + (size_t)b * nsb; Cd.forced_idx = nullptr;
      if (cdef_on) CK(launch_cdef(Cd, cnt, e->stream));
      e->kernel_launches += cdef_on ? 2 : 1;
      CK(cudaEventRecord(ev[3], e->stream));
      if (e->lr_on) {
        LrLaunch R;
        R.g = g; R.bit_depth = bd;
        for (int p = 0; p < 3; p++) {
          R.lr_type[p] = p ? AV1B_RESTORE_NONE : AV1B_RESTORE_SWITCHABLE;
          R.unit_size[p] = 64 >> (p > 0); R.unit_rows[p] = e->lr_rows; R.unit_cols[p] = e->lr_cols;
          R.cdef[p] = cdef_on ? rec[p] : deb[p]; R.deb[p] = deb[p]; R.out[p] = fin[p]; R.plane_elems[p] = e->plane_elems[p]; R.units[p] = nullptr;
        }
        Av1bLrUnit* units = s.d_lr_units + (size_t)b * e->lr_n;
        R.src_y = src[0]; R.cand = e->lr_cand; R.sse = e->d_lr_sse;
        const long long aq = acq;
        CK(launch_lr_search(R, cnt, (aq * aq * 5) >> 8, units, e->stream));   // rate of a unit's parameters in squared-error units
        R.units[0] = units;
        CK(launch_lr(R, cnt, e->stream));
        e->kernel_launches += 3;
      }
    } else {
      CK(cudaEventRecord(ev[2], e->stream));
      CK(cudaEventRecord(ev[3], e->stream));
    }
    CK(cudaEventRecord(ev[4], e->stream));
  }
  e->chunk_pos = gop_pos[n - 1] + 1;
  if (e->quality_on) {
    CK(launch_quality(g, bd, e->d_fin[0] + e->plane_elems[0], in.d_src[0], e->plane_elems[0], s.d_quality, n, e->stream));
    e->kernel_launches += 1;
  }
  if (e->keep) {
    // debug mode keeps every reconstruction: download it before the ring is shifted
    for (int p = 0; p < 3; p++)
      CK(cudaMemcpyAsync(s.h_rec[p], e->d_fin[p] + e->plane_elems[p], e->plane_elems[p] * n * 2, cudaMemcpyDeviceToHost, e->stream));
  }
  // the last key / anchor picture of the batch becomes ring entry 0: the reference the next batch starts from
  if (last_ref > 0)
    for (int p = 0; p < 3; p++)
      CK(cudaMemcpyAsync(e->d_fin[p], e->d_fin[p] + e->plane_elems[p] * last_ref, e->plane_elems[p] * 2, cudaMemcpyDeviceToDevice, e->stream));
  // ---- inter frames of the batch: tokens for the host range coder ----
  s.has_tokens = e->token_path && any_inter;
  CK(cudaEventRecord(s.ev_tok0, e->stream));
  if (s.has_tokens) {
    TokLaunch& T = s.tl;
    T.g = e->g_inter; T.n_frames = n; T.inter_mask = 0;
    for (int b = 0; b < n; b++) if (!s.is_key[b]) T.inter_mask |= (uint64_t)1 << b;
    T.cdef_bits = e->loop_filters ? e->fp_inter.cdef_bits : 0;
    T.nocdef_mask = 0;
    for (int b = 0; b < n; b++) if (kind_params(e, s.kind[b]).cdef_bits == 0) T.nocdef_mask |= (uint64_t)1 << b;
    T.blocks = s.d_blocks; T.cdef_idx = s.d_cdef_idx;
    for (int p = 0; p < 3; p++) { T.digest[p] = s.d_digest[p]; T.coef[p] = s.d_coef[p]; T.plane_elems[p] = e->plane_elems[p]; }
    T.map_elems = e->map_elems;
    T.mode_cls = s.d_mode_cls; T.blk_count = s.d_blk_count; T.sb_off = s.d_sb_off;
    T.sb_of_order = e->d_sb_of_order; T.tile_of_sb = e->d_tile_of_sb;
    T.tokens = s.d_tokens; T.cap = (uint32_t)s.tok_cap;
    T.lr_units = e->lr_on ? s.d_lr_units : nullptr; T.lr_rows = e->lr_rows; T.lr_cols = e->lr_cols;
    CK(launch_tok_count(T, e->stream));
    CK(launch_tok_emit(T, e->stream));
    e->kernel_launches += 4;
  }
  CK(cudaEventRecord(s.ev_tok1, e->stream));
  CK(cudaEventRecord(s.ev_k1, e->stream));
  const bool rc = s.has_tokens && e->rc_on;
  if (rc) {
    // the range coder of this batch runs beside the kernels of the next one
    RcLaunch& R = s.rl;
    R.n_frames = n; R.n_tiles = e->g_inter.tile_cols * e->g_inter.tile_rows; R.nsb = (int)nsb; R.inter_mask = s.tl.inter_mask;
    R.tokens = s.d_tokens; R.tok_cap = (uint32_t)s.tok_cap; R.sb_off = s.d_sb_off; R.tile_first_k = e->d_tile_first_k; R.cdf_init = e->d_cdf_init;
    R.cdf_init_alt = e->d_cdf_init_alt; R.alt_mask = 0;
    for (int b = 0; b < n; b++) if (s.kind[b] == 2) R.alt_mask |= (uint64_t)1 << b;
    R.region = s.d_rc_region; R.tile_len = s.d_rc_len; R.bytes = s.d_rc_bytes; R.cap_bytes = (uint32_t)s.rc_cap; R.overflow = e->d_rc_overflow;
    CK(cudaStreamWaitEvent(s.s_rc, s.ev_k1, 0));
    CK(cudaEventRecord(s.ev_rc0, s.s_rc));
    CK(launch_rc(R, s.s_rc));
    CK(cudaEventRecord(s.ev_rc1, s.s_rc));
    e->kernel_launches += 3;
  }
  // symbol streams go home on the output copy stream while the compute stream starts the next batch:
  // token offsets for the inter frames (the token list itself follows once its size is known), levels and
  // block info only for the frames the block-walking tile writer codes (key frames; everything in debug mode)
  CK(cudaStreamWaitEvent(e->s_out, s.ev_k1, 0));
  if (e->lr_on) {
    CK(cudaMemcpyAsync(s.h_lr_units, s.d_lr_units, e->lr_n * n * sizeof(Av1bLrUnit), cudaMemcpyDeviceToHost, e->s_out));
    e->d2h_bytes += (int64_t)(e->lr_n * n * sizeof(Av1bLrUnit));
  }
  if (e->quality_on) {
    CK(cudaMemcpyAsync(s.h_quality, s.d_quality, sizeof(QualityAcc) * n, cudaMemcpyDeviceToHost, e->s_out));
    e->d2h_bytes += (int64_t)(sizeof(QualityAcc) * n);
  }
  s.tok_fetched = 0;
  if (s.has_tokens) {
    CK(cudaMemcpyAsync(s.h_sb_off, s.d_sb_off, (nsb * n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->s_out));
    e->d2h_bytes += (int64_t)((nsb * n + 1) * sizeof(uint32_t));
    // the token total is only known once the offsets are home: fetch as many tokens as the last batch had (+25 %)
    // right away, finish() fetches what is missing
    s.tok_fetched = std::min(s.tok_cap, e->tok_guess + e->tok_guess / 4 + 65536);
    if (e->tok_guess == 0 || rc) s.tok_fetched = 0;   // with the device range coder the tokens stay on the device
    if (s.tok_fetched) CK(cudaMemcpyAsync(s.h_tokens, s.d_tokens, s.tok_fetched * 4, cudaMemcpyDeviceToHost, e->s_out));
    e->d2h_bytes += (int64_t)s.tok_fetched * 4;
  }
  for (int b = 0; b < n; b++) {
    if (!(e->keep || s.is_key[b] || !e->token_path)) continue;
    for (int p = 0; p < 3; p++) {
      if (!s.h_coef[p]) CK(cudaMallocHost(&s.h_coef[p], e->plane_elems[p] * (size_t)e->batch * 2));
      CK(cudaMemcpyAsync(s.h_coef[p] + (size_t)b * e->plane_elems[p], s.d_coef[p] + (size_t)b * e->plane_elems[p], e->plane_elems[p] * 2, cudaMemcpyDeviceToHost, e->s_out));
    }
    CK(cudaMemcpyAsync(s.h_cdef_idx + (size_t)b * nsb, s.d_cdef_idx + (size_t)b * nsb, nsb, cudaMemcpyDeviceToHost, e->s_out));
    CK(cudaMemcpyAsync(s.h_blocks + (size_t)b * e->map_elems, s.d_blocks + (size_t)b * e->map_elems, e->map_elems * sizeof(Av1bBlockInfo), cudaMemcpyDeviceToHost, e->s_out));
    e->d2h_bytes += (int64_t)((e->plane_elems[0] + e->plane_elems[1] + e->plane_elems[2]) * 2 + nsb + e->map_elems * sizeof(Av1bBlockInfo));
  }
  s.rc_fetched = 0;
  if (rc) {
    const size_t nt = (size_t)s.rl.n_tiles * n + 1;
    CK(cudaStreamWaitEvent(e->s_out, s.ev_rc1, 0));
    CK(cudaMemcpyAsync(s.h_rc_len, s.d_rc_len, nt * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->s_out));
    CK(cudaMemcpyAsync(s.h_rc_len + (size_t)s.rl.n_tiles * e->batch + 1, e->d_rc_overflow, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->s_out));
    e->d2h_bytes += (int64_t)((nt + 1) * sizeof(uint32_t));
    s.rc_fetched = e->rc_guess ? std::min(s.rc_cap, e->rc_guess + e->rc_guess / 4 + 65536) : 0;
    if (s.rc_fetched) CK(cudaMemcpyAsync(s.h_rc_bytes, s.d_rc_bytes, s.rc_fetched, cudaMemcpyDeviceToHost, e->s_out));
    e->d2h_bytes += (int64_t)s.rc_fetched;
  }
  CK(cudaEventRecord(s.ev_d2h, e->s_out));
  return AV1B_OK;
}

// wait for the slot's download, entropy-code its frames on the host pool, hand out packets in order
static int finish(av1b_encoder* e, Slot& s, bool staged, av1b_packet_cb out_cb, av1b_progress_cb prog_cb, void* user,
                  int64_t total_frames, std::chrono::steady_clock::time_point t_start) {
  const Av1bGeom& g = e->g;
  CK(cudaEventSynchronize(s.ev_d2h));
  float ms = 0;
  if (staged) collect_h2d(e, s);
  cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1); e->t_kernel_ms += ms;
  cudaEventElapsedTime(&ms, s.ev_k0, s.ev_me); e->t_me_ms += ms;
  if (e->mctf_on && !e->intra_only) { cudaEventElapsedTime(&ms, s.ev_tf0, s.ev_tf1); e->t_mctf_ms += ms; }
  cudaEventElapsedTime(&ms, s.ev_k1, s.ev_d2h); e->t_d2h_ms += ms;
  const int n = s.n_frames;
  for (size_t gi = 0; gi < s.groups.size(); gi++) {
    cudaEvent_t* ev = &s.ev_frame[gi * 5];
    cudaEventElapsedTime(&ms, ev[0], ev[1]);
    if (s.is_key[s.groups[gi].first]) e->t_intra_ms += ms; else e->t_inter_ms += ms;
    if (e->loop_filters) {
      cudaEventElapsedTime(&ms, ev[1], ev[2]); e->t_deblock_ms += ms;
      cudaEventElapsedTime(&ms, ev[2], ev[3]); e->t_cdef_ms += ms;
      cudaEventElapsedTime(&ms, ev[3], ev[4]); e->t_lr_ms += ms;
    }
  }
  cudaEventElapsedTime(&ms, s.ev_tok0, s.ev_tok1); e->t_tok_ms += ms;
  const size_t nsb = (size_t)g.sb_rows * g.sb_cols;
  const bool rc = s.has_tokens && e->rc_on;
  if (s.has_tokens) {
    // the batch's token total is known now: make room if needed (then the tokens are written again), fetch them
    size_t total = s.h_sb_off[nsb * n];
    bool redo = false;
    if (total > s.tok_cap) {
      const size_t cap = total + total / 4;
      cudaFree(s.d_tokens); cudaFreeHost(s.h_tokens); s.d_tokens = nullptr; s.h_tokens = nullptr; s.tok_cap = 0;
      if (cudaMalloc(&s.d_tokens, cap * 4) != cudaSuccess || cudaMallocHost(&s.h_tokens, cap * 4) != cudaSuccess) {
        set_error("token buffers (%zu tokens): out of memory", cap); return AV1B_ERR_NOMEM;
      }
      s.tok_cap = cap;
      s.tl.tokens = s.d_tokens; s.tl.cap = (uint32_t)cap;
      if (e->rc_on) {
        cudaFree(s.d_rc_region); s.d_rc_region = nullptr;
        if (cudaMalloc(&s.d_rc_region, 2 * cap + 64 * (size_t)s.rl.n_tiles * e->batch) != cudaSuccess) { set_error("range coder scratch: out of memory"); return AV1B_ERR_NOMEM; }
        s.rl.tokens = s.d_tokens; s.rl.tok_cap = (uint32_t)cap; s.rl.region = s.d_rc_region;
      }
      CK(launch_tok_emit(s.tl, e->stream));
      CK(cudaStreamSynchronize(e->stream));
      s.tok_fetched = 0;
      redo = true;
    }
    if (rc) {
      const size_t nt = (size_t)s.rl.n_tiles * n;
      for (int attempt = 0; attempt < 3; attempt++) {
        if (redo) {
          CK(launch_rc(s.rl, e->stream));
          CK(cudaMemcpyAsync(s.h_rc_len, s.d_rc_len, (nt + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
          CK(cudaStreamSynchronize(e->stream));
          s.rc_fetched = 0;
        }
        const size_t bytes = s.h_rc_len[nt];
        if (bytes <= s.rc_cap) break;
        // payloads larger than the byte buffers: grow them and code again
        const size_t cap = bytes + bytes / 4;
        cudaFree(s.d_rc_bytes); cudaFreeHost(s.h_rc_bytes); s.d_rc_bytes = nullptr; s.h_rc_bytes = nullptr; s.rc_cap = 0;
        if (cudaMalloc(&s.d_rc_bytes, cap) != cudaSuccess || cudaMallocHost(&s.h_rc_bytes, cap) != cudaSuccess) { set_error("payload buffers: out of memory"); return AV1B_ERR_NOMEM; }
        s.rc_cap = cap; s.rl.bytes = s.d_rc_bytes; s.rl.cap_bytes = (uint32_t)cap;
        redo = true;
      }
      const size_t bytes = s.h_rc_len[nt];
      if (bytes > s.rc_fetched) {
        const auto tc0 = std::chrono::steady_clock::now();
        CK(cudaMemcpyAsync(s.h_rc_bytes + s.rc_fetched, s.d_rc_bytes + s.rc_fetched, bytes - s.rc_fetched, cudaMemcpyDeviceToHost, e->s_tok));
        CK(cudaStreamSynchronize(e->s_tok));
        e->d2h_bytes += (int64_t)(bytes - s.rc_fetched);
        e->t_d2h_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count();
      }
      e->rc_guess = bytes;
      if (s.h_rc_len[(size_t)s.rl.n_tiles * e->batch + 1]) { set_error("range coder: a tile outgrew its region (2 bytes per token + 64)"); return AV1B_ERR_INTERNAL; }
      cudaEventElapsedTime(&ms, s.ev_rc0, s.ev_rc1); e->t_rc_ms += ms;
    } else if (total > s.tok_fetched) {
      const auto tc0 = std::chrono::steady_clock::now();
      CK(cudaMemcpyAsync(s.h_tokens + s.tok_fetched, s.d_tokens + s.tok_fetched, (total - s.tok_fetched) * 4, cudaMemcpyDeviceToHost, e->s_tok));
      CK(cudaStreamSynchronize(e->s_tok));
      e->d2h_bytes += (int64_t)(total - s.tok_fetched) * 4;
      e->t_d2h_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count();
    }
    e->tok_guess = total;
    e->n_tokens += (int64_t)total;
  }
  const auto tp0 = std::chrono::steady_clock::now();
  std::vector<Av1bFrameSyms> sy(n);
  std::vector<FramePack> packs(n);
  std::vector<std::pair<int, int>> tasks;   // (frame, tile)
  for (int b = 0; b < n; b++) {
    const Av1bGeom& gb = s.is_key[b] ? e->g : e->g_inter;
    memset(&sy[b], 0, sizeof(Av1bFrameSyms));
    sy[b].blocks = s.h_blocks + (size_t)b * e->map_elems;
    for (int p = 0; p < 3; p++) { sy[b].coef[p] = s.h_coef[p] ? s.h_coef[p] + (size_t)b * e->plane_elems[p] : nullptr; sy[b].coef_stride[p] = g.stride[p]; }
    sy[b].cdef_idx = s.h_cdef_idx + (size_t)b * g.sb_rows * g.sb_cols;
    if (e->lr_on) { sy[b].lr_units[0] = s.h_lr_units + (size_t)b * e->lr_n; sy[b].lr_unit_rows[0] = e->lr_rows; sy[b].lr_unit_cols[0] = e->lr_cols; }
    Av1bFrameParams fph = kind_params(e, s.kind[b]);
    if (e->grain_scaling > 0) {
      const uint32_t fi = (uint32_t)(s.first_index + b);
      fph.grain_scaling = e->grain_scaling;
      fph.grain_seed = (int32_t)(((fi + 1) * 2654435761u) >> 16);   // another seed for every frame
    }
    pack_frame_header(e->seq, fph, gb, packs[b]);
    for (int t = 0; t < gb.tile_cols * gb.tile_rows; t++) tasks.emplace_back(b, t);
  }
  if (s.has_tokens && !rc) {
    // longest tiles first: the pool's tail is then a short tile, not a long one
    auto cost = [&](const std::pair<int, int>& x) -> uint32_t {
      if (s.is_key[x.first]) return 0xFFFFFFFFu;
      const uint32_t* off = s.h_sb_off + (size_t)x.first * nsb;
      return off[e->tile_first_k[x.second + 1]] - off[e->tile_first_k[x.second]];
    };
    std::stable_sort(tasks.begin(), tasks.end(), [&](const std::pair<int, int>& a, const std::pair<int, int>& b) { return cost(a) > cost(b); });
  }
  e->pool->parallel_for((int)tasks.size(), [&](int t) {
    const int b = tasks[t].first, tile = tasks[t].second;
    if (!s.is_key[b] && rc) {
      const uint32_t* o = s.h_rc_len + (size_t)b * s.rl.n_tiles + tile;
      packs[b].tiles[tile].assign(s.h_rc_bytes + o[0], s.h_rc_bytes + o[1]);
    } else if (!s.is_key[b] && s.has_tokens) {
      const uint32_t* off = s.h_sb_off + (size_t)b * nsb;
      const uint32_t t0 = off[e->tile_first_k[tile]];
      // the entry after a frame's last superblock is the next frame's first (or the batch total)
      const uint32_t t1 = off[e->tile_first_k[tile + 1]];
      pack_tile_tokens(kind_params(e, s.kind[b]), s.h_tokens + t0, t1 - t0, packs[b].tiles[tile]);
    } else {
      pack_tile(e->seq, kind_params(e, s.kind[b]), s.is_key[b] ? e->g : e->g_inter, sy[b], tile, packs[b].tiles[tile]);
    }
  });
  std::vector<uint8_t> tu;
  for (int b = 0; b < n; b++) {
    tu.clear();
    write_temporal_delimiter(tu);
    if (s.is_key[b]) write_sequence_header(e->seq, tu);
    assemble_frame(packs[b], tu);
    if (e->keep) {
      e->kept.emplace_back();
      KeptFrame& k = e->kept.back();
      for (int p = 0; p < 3; p++) {
        k.rec[p].assign(s.h_rec[p] + (size_t)b * e->plane_elems[p], s.h_rec[p] + (size_t)(b + 1) * e->plane_elems[p]);
        k.coef[p].assign(s.h_coef[p] + (size_t)b * e->plane_elems[p], s.h_coef[p] + (size_t)(b + 1) * e->plane_elems[p]);
      }
      k.blocks.assign(sy[b].blocks, sy[b].blocks + e->map_elems);
      k.cdef_idx.assign(sy[b].cdef_idx, sy[b].cdef_idx + (size_t)g.sb_rows * g.sb_cols);
      if (e->lr_on) k.lr_units.assign(sy[b].lr_units[0], sy[b].lr_units[0] + e->lr_n);
      k.is_key = s.is_key[b]; k.kind = s.kind[b];
    }
    e->bytes_out += (int64_t)tu.size();
    if (out_cb && out_cb(user, tu.data(), tu.size(), s.first_index + b, s.is_key[b])) { set_error("packet callback aborted"); return AV1B_ERR_CALLBACK; }
  }
  e->frames_done += n;
  if (e->quality_on)
    for (int b = 0; b < n; b++) {
      const QualityAcc& qa = s.h_quality[b];
      const double px = (double)qa.blocks * 64, peak = (double)((1 << e->cfg.bit_depth) - 1);
      if (px > 0) {
        e->q_psnr_sum += qa.sse ? 10.0 * log10(peak * peak * px / (double)qa.sse) : 100.0;
        e->q_ssim += qa.ssim_sum / qa.blocks; e->q_frames += 1;
      }
    }
  const auto tp1 = std::chrono::steady_clock::now();
  e->t_pack_ms += std::chrono::duration<double, std::milli>(tp1 - tp0).count();
  if (prog_cb) {
    const double el = std::chrono::duration<double>(tp1 - t_start).count();
    prog_cb(user, e->frames_done, total_frames, el > 0 ? e->frames_done / el : 0.0);
  }
  return AV1B_OK;
}

static void reset_stats(av1b_encoder* e) {
  e->kept.clear();
  e->t_h2d_ms = e->t_kernel_ms = e->t_intra_ms = e->t_inter_ms = e->t_me_ms = e->t_d2h_ms = e->t_pack_ms = 0;
  e->q_psnr_sum = e->q_ssim = e->q_frames = 0;
  e->t_deblock_ms = e->t_cdef_ms = e->t_tok_ms = 0; e->t_lr_ms = 0; e->t_rc_ms = 0; e->n_tokens = 0; e->d2h_bytes = 0; e->t_mctf_ms = 0; e->mctf_frames = 0;
  e->kernel_launches = e->intra_launches = e->inter_launches = e->frames_done = e->bytes_out = e->key_frames = e->staged_direct = 0;
  e->scene_cuts = 0;
}

extern "C" {

int av1b_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void av1b_config_default(av1b_config* c) {
  memset(c, 0, sizeof(*c));
  c->bit_depth = 10; c->fps_num = 30; c->fps_den = 1;
  c->crf = 30; c->preset = 6; c->keyint = 240; c->lookahead = -1;
  c->tile_cols_log2 = -1; c->tile_rows_log2 = -1;
  c->qm_min = 8; c->qm_max = 15;   // SVT-AV1's defaults; only read with enable_qm
}

static void set_structure(av1b_encoder* e, int gop_period);

int av1b_encoder_create(const av1b_config* cfg, av1b_encoder** out) {
  if (!cfg || !out) { set_error("null argument"); return AV1B_ERR_INVALID; }
  *out = nullptr;
  if (cfg->bit_depth != 8 && cfg->bit_depth != 10) { set_error("bit_depth must be 8 or 10"); return AV1B_ERR_INVALID; }
  if (cfg->crf < 0 || cfg->crf > 63) { set_error("crf out of range 0..63"); return AV1B_ERR_INVALID; }
  if (cfg->enable_qm && (cfg->qm_min < 0 || cfg->qm_max > 15 || cfg->qm_min > cfg->qm_max)) {
    set_error("qm_min / qm_max must satisfy 0 <= qm_min <= qm_max <= 15"); return AV1B_ERR_INVALID;
  }
  // the coded frame: the source padded to multiples of 8 by edge replication (stage()); the frame headers then carry the
  // source size as render_size (1920x804 and the like: scope crops of 1080p / 2160p material)
  const int coded_w = (cfg->width + 7) & ~7, coded_h = (cfg->height + 7) & ~7;
  Av1bGeom probe;
  if (cfg->width < 16 || cfg->height < 16 || av1b_geom_init(&probe, coded_w, coded_h, 0, 0)) {
    set_error("unsupported frame size %dx%d (16..8192 x 16..4352)", cfg->width, cfg->height);
    return AV1B_ERR_INVALID;
  }
  int ndev = av1b_device_count();
  if (ndev <= 0) { set_error("no CUDA device visible (av1b200 has no CPU fallback)"); return AV1B_ERR_NO_DEVICE; }
  if (cfg->device_id < 0 || cfg->device_id >= ndev) { set_error("device_id %d out of range", cfg->device_id); return AV1B_ERR_INVALID; }
  av1b_encoder* e = new av1b_encoder();
  e->cfg = *cfg;
  e->src_w = cfg->width; e->src_h = cfg->height;
  e->padded_src = coded_w != cfg->width || coded_h != cfg->height;
  // key-frame tiles: auto = 2x2 superblocks per tile (every tile is one serial chain of the closed-loop intra
  // kernel, so many small tiles = many parallel chains; key frames are rare, the extra tile overhead is cheap)
  int tcl = cfg->tile_cols_log2, trl = cfg->tile_rows_log2;
  if (tcl < 0) tcl = av1b_tile_log2(2, probe.sb_cols);
  if (trl < 0) trl = av1b_tile_log2(2, probe.sb_rows);
  av1b_geom_init(&e->g, coded_w, coded_h, tcl, trl);
  {
    // tiles of inter frames are the unit of entropy-coder parallelism.  The host coder wants few large ones (longer
    // CDF adaptation): about 12x12 superblocks.  The device coder walks a tile with one warp, a serial chain of about
    // 100 ns per symbol, so its latency is that of the largest tile: about 6x6 superblocks (reserved[7] overrides).
    // Where the range coder runs (reserved[5] = 0: automatic): a 150-frame 4K chunk at CRF 30 is about 230 thread-ms of host
    // range coding against 75 ms of kernels, so four host threads for this GPU keep up with it (8 GPUs on a 32-core box:
    // 14156 fps with the host coder, 13041 fps with the device coder whose single-warp chains then bound the step,
    // profiles/r02j_scale_n8_*.json); with fewer threads the host would throttle the GPU and the device codes.
    const int ht = cfg->host_threads > 0 ? cfg->host_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    e->rc_on = cfg->reserved[5] == 4 || (cfg->reserved[5] == 0 && ht < 4);
    e->n_slots = (e->rc_on && e->token_path && cfg->reserved[3] == 0) ? 4 : 3;
    // (4x4 up to 1080p, where a batch is short and the frame has few tiles)
    const int tsb = cfg->reserved[7] > 0 ? cfg->reserved[7] : (e->rc_on ? (probe.sb_cols * probe.sb_rows <= 600 ? 4 : 6) : 12);
    const int itc = cfg->tile_cols_log2 >= 0 ? cfg->tile_cols_log2 : av1b_tile_log2(tsb, probe.sb_cols);
    const int itr = cfg->tile_rows_log2 >= 0 ? cfg->tile_rows_log2 : av1b_tile_log2(tsb, probe.sb_rows);
    av1b_geom_init(&e->g_inter, coded_w, coded_h, itc, itr);
  }
  e->seq.width = coded_w; e->seq.height = coded_h; e->seq.bit_depth = cfg->bit_depth;
  e->seq.render_width = e->padded_src ? cfg->width : 0; e->seq.render_height = e->padded_src ? cfg->height : 0;
  e->loop_filters = cfg->reserved[2] == 0;     // reserved[2] = 1 switches the in-loop filters off (tests)
  // --preset <= 5 (the daemon passes 3, av1an.rs:14) adds loop restoration with a per-unit decision; reserved[6] = 1 keeps it off
  e->lr_on = e->loop_filters && cfg->preset <= 5 && cfg->reserved[6] == 0;
  e->seq.enable_cdef = e->loop_filters ? 1 : 0; e->seq.enable_restoration = e->lr_on ? 1 : 0;
  e->seq.fps_num = cfg->fps_num; e->seq.fps_den = cfg->fps_den; e->seq.color_hdr = cfg->hdr;
  e->seq.film_grain_present = cfg->film_grain > 0 ? 1 : 0;
  e->base_q_idx = av1t_quantizer_to_qindex[cfg->crf];
  if (e->base_q_idx < 1) e->base_q_idx = 1;   // lossless (qindex 0) is not supported
  e->base_q_idx_key = cfg->reserved[3] ? e->base_q_idx : std::max(1, e->base_q_idx * 3 / 4);
  // one-level hierarchy: every gop_period-th frame is an anchor (the only inter frames that become references, coded
  // a little finer), the frames between two anchors predict from the last anchor at a much coarser quantiser;
  // config.gop_period == 0: chosen per chunk from the noise level of the source (begin_chunk)
  if (cfg->gop_period > 16) { set_error("gop_period must be <= 16"); delete e; return AV1B_ERR_INVALID; }
  e->q_nominal = e->base_q_idx;
  e->gop_auto = cfg->gop_period == 0 && cfg->reserved[3] == 0;
  e->mctf_cfg = cfg->tune[2] == 0 && cfg->reserved[3] == 0;
  e->quality_on = cfg->tune[3] != 0;
  e->me_smooth = cfg->tune[0] == 0;
  e->scene_on = cfg->tune[4] == 0;
  e->key_var_part = cfg->tune[1] == 0 && cfg->reserved[1] == 0;
  e->blk_log2 = cfg->reserved[1] ? cfg->reserved[1] : 4;
  if (e->blk_log2 < 3 || e->blk_log2 > 6) { set_error("block log2 must be 3..6"); delete e; return AV1B_ERR_INVALID; }
  e->keep = cfg->reserved[0] != 0;
  e->token_path = cfg->reserved[5] == 0 || cfg->reserved[5] == 3 || cfg->reserved[5] == 4;
  e->legacy_pack_levels = (cfg->reserved[5] == 2 && !e->keep) ? 1 : 0;
  e->intra_only = cfg->reserved[3] != 0;      // reserved[3] = 1: every frame is a key frame
  e->keyint = cfg->keyint > 0 ? cfg->keyint : 240;
  av1b_select_frame_params(cfg->bit_depth, e->base_q_idx_key, AV1B_KEY_FRAME, e->loop_filters ? 1 : 0, &e->fp_key);
  set_qm_levels(e, &e->fp_key);
  if (e->lr_on) {
    for (Av1bFrameParams* f : {&e->fp_key}) { f->lr_type[0] = AV1B_RESTORE_SWITCHABLE; f->lr_type[1] = f->lr_type[2] = AV1B_RESTORE_NONE; f->lr_unit_shift = 0; f->lr_uv_shift = 0; }
    e->lr_rows = std::max((coded_h + 32) / 64, 1); e->lr_cols = std::max((coded_w + 32) / 64, 1);
    e->lr_n = (size_t)e->lr_rows * e->lr_cols;
    memset(&e->lr_cand, 0, sizeof(e->lr_cand));
    // a mild symmetric smoother and the radius-1 self-guided filter with a small weight: what the decision picks from
    e->lr_cand.wiener_v[2] = 8; e->lr_cand.wiener_h[2] = 8; e->lr_cand.sgr_set = 12; e->lr_cand.sgr_xqd[0] = 0; e->lr_cand.sgr_xqd[1] = 95;
  }
  e->fp_key.tile_cols_log2 = e->g.tile_cols_log2; e->fp_key.tile_rows_log2 = e->g.tile_rows_log2;
  set_structure(e, e->intra_only ? 1 : (cfg->gop_period > 0 ? cfg->gop_period : kDefaultGopPeriod));
  e->batch = cfg->frames_in_flight > 0 ? cfg->frames_in_flight : 8;
  if (e->batch > 64) { set_error("frames_in_flight must be <= 64"); delete e; return AV1B_ERR_INVALID; }
  e->host_threads = cfg->host_threads > 0 ? cfg->host_threads : (int)std::max(1u, std::thread::hardware_concurrency());
  if (cudaSetDevice(cfg->device_id) != cudaSuccess) { set_error("cudaSetDevice failed"); delete e; return AV1B_ERR_CUDA; }
  cudaError_t err = cudaSuccess;
  auto A = [&](cudaError_t r) { if (err == cudaSuccess && r != cudaSuccess) err = r; };
  A(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  A(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
  A(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
  A(cudaStreamCreateWithFlags(&e->s_tok, cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    for (int si = 0; si < e->n_slots; si++) A(cudaStreamCreateWithPriority(&e->slot[si].s_rc, cudaStreamNonBlocking, hi));   // its few warps should start promptly
  }
  e->map_elems = (size_t)e->g.w8 * e->g.h8;
  const size_t nsb = (size_t)e->g.sb_rows * e->g.sb_cols;
  for (int p = 0; p < 3; p++) e->plane_elems[p] = (size_t)e->g.stride[p] * e->g.rows[p];
  const int F = e->batch;
  for (int si = 0; si < e->n_slots; si++) {
    Slot& s = e->slot[si];
    for (cudaEvent_t* ev : {&s.ev_h2d, &s.ev_k0, &s.ev_me, &s.ev_k1, &s.ev_d2h, &s.ev_src, &s.ev_tok0, &s.ev_tok1, &s.ev_rc0, &s.ev_rc1, &s.ev_tf0, &s.ev_tf1}) A(cudaEventCreate(ev));
    s.ev_frame.assign((size_t)F * 5, nullptr);
    for (auto& ev : s.ev_frame) A(cudaEventCreate(&ev));
    for (int p = 0; p < 3; p++) {
      const size_t n = e->plane_elems[p] * F;
      A(cudaMalloc(&s.d_src[p], n * 2)); A(cudaMalloc(&s.d_coef[p], n * 2));
      // the page-locked staging copy of the sources (pageable callers) and the page-locked mirror of the levels (key frames,
      // debug mode) are allocated when they are first needed: 400 MB per slot at 4K that a job with page-locked sources and
      // one key frame per chunk hardly touches, and pinning memory is what makes encoder start-up slow
      if (e->keep) A(cudaMallocHost(&s.h_rec[p], n * 2));
      if (err == cudaSuccess) { A(cudaMemset(s.d_src[p], 0, n * 2)); A(cudaMemset(s.d_coef[p], 0, n * 2)); }
    }
    A(cudaMalloc(&s.d_score, F * sizeof(uint32_t))); A(cudaMallocHost(&s.h_score, F * sizeof(uint32_t))); A(cudaEventCreate(&s.ev_score));
    A(cudaMallocHost(&s.h_blocks, e->map_elems * F * sizeof(Av1bBlockInfo)));
    A(cudaMallocHost(&s.h_cdef_idx, nsb * F));
    A(cudaMalloc(&s.d_blocks, e->map_elems * F * sizeof(Av1bBlockInfo)));
    A(cudaMalloc(&s.d_cdef_idx, nsb * F));
    if (cfg->tune[3]) { A(cudaMalloc(&s.d_quality, sizeof(QualityAcc) * F)); A(cudaMallocHost(&s.h_quality, sizeof(QualityAcc) * F)); }
    if (err == cudaSuccess) { A(cudaMemset(s.d_blocks, 0, e->map_elems * F * sizeof(Av1bBlockInfo))); A(cudaMemset(s.d_cdef_idx, 0, nsb * F)); }
    if (e->lr_on) { A(cudaMalloc(&s.d_lr_units, e->lr_n * F * sizeof(Av1bLrUnit))); A(cudaMallocHost(&s.h_lr_units, e->lr_n * F * sizeof(Av1bLrUnit))); }
    if (e->token_path && !e->intra_only) {
      for (int p = 0; p < 3; p++) A(cudaMalloc(&s.d_digest[p], e->plane_elems[p] * F * 2));
      A(cudaMalloc(&s.d_mode_cls, e->map_elems * F));
      A(cudaMalloc(&s.d_blk_count, e->map_elems * F * sizeof(uint32_t)));
      A(cudaMalloc(&s.d_sb_off, (nsb * F + 1) * sizeof(uint32_t)));
      A(cudaMallocHost(&s.h_sb_off, (nsb * F + 1) * sizeof(uint32_t)));
      // room for one token per four luma samples (several times what CRF 30 produces); grows on demand
      s.tok_cap = (size_t)coded_w * coded_h / 4 * F;
      A(cudaMalloc(&s.d_tokens, s.tok_cap * 4));
      A(cudaMallocHost(&s.h_tokens, s.tok_cap * 4));
      if (e->rc_on) {
        const size_t nt = (size_t)e->g_inter.tile_cols * e->g_inter.tile_rows * F;
        A(cudaMalloc(&s.d_rc_region, 2 * s.tok_cap + 64 * nt));
        A(cudaMalloc(&s.d_rc_len, (nt + 1) * sizeof(uint32_t)));
        A(cudaMallocHost(&s.h_rc_len, (nt + 2) * sizeof(uint32_t)));   // + the overflow flag
        s.rc_cap = s.tok_cap / 2;   // bytes; grows on demand
        A(cudaMalloc(&s.d_rc_bytes, s.rc_cap));
        A(cudaMallocHost(&s.h_rc_bytes, s.rc_cap));
      }
    }
  }
  if (e->token_path && !e->intra_only) {
    // coding order of the superblocks of an inter frame: tile by tile, raster order inside a tile
    const Av1bGeom& gi = e->g_inter;
    std::vector<uint32_t> order; std::vector<uint16_t> tile_of(nsb, 0);
    e->tile_first_k.clear();
    for (int tr = 0; tr < gi.tile_rows; tr++)
      for (int tc = 0; tc < gi.tile_cols; tc++) {
        e->tile_first_k.push_back((uint32_t)order.size());
        for (int r = gi.tile_row_start_sb[tr]; r < gi.tile_row_start_sb[tr + 1]; r++)
          for (int c = gi.tile_col_start_sb[tc]; c < gi.tile_col_start_sb[tc + 1]; c++) {
            order.push_back((uint32_t)(r * gi.sb_cols + c));
            tile_of[(size_t)r * gi.sb_cols + c] = (uint16_t)(tr * gi.tile_cols + tc);
          }
      }
    e->tile_first_k.push_back((uint32_t)order.size());
    A(cudaMalloc(&e->d_sb_of_order, nsb * sizeof(uint32_t)));
    A(cudaMalloc(&e->d_tile_of_sb, nsb * sizeof(uint16_t)));
    if (err == cudaSuccess) {
      A(cudaMemcpy(e->d_sb_of_order, order.data(), nsb * sizeof(uint32_t), cudaMemcpyHostToDevice));
      A(cudaMemcpy(e->d_tile_of_sb, tile_of.data(), nsb * sizeof(uint16_t), cudaMemcpyHostToDevice));
    }
    if (e->rc_on) {
      std::vector<uint8_t> img(tile_cdfs_size()), img2(tile_cdfs_size());
      tile_cdfs_default(e->base_q_idx, img.data());
      tile_cdfs_default(e->base_q_idx_nonref, img2.data());
      A(cudaMalloc(&e->d_cdf_init, img.size()));
      A(cudaMalloc(&e->d_cdf_init_alt, img2.size()));
      if (err == cudaSuccess) A(cudaMemcpy(e->d_cdf_init_alt, img2.data(), img2.size(), cudaMemcpyHostToDevice));
      A(cudaMalloc(&e->d_tile_first_k, e->tile_first_k.size() * sizeof(uint32_t)));
      A(cudaMalloc(&e->d_rc_overflow, sizeof(uint32_t)));
      if (err == cudaSuccess) {
        A(cudaMemcpy(e->d_cdf_init, img.data(), img.size(), cudaMemcpyHostToDevice));
        A(cudaMemcpy(e->d_tile_first_k, e->tile_first_k.data(), e->tile_first_k.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        A(cudaMemset(e->d_rc_overflow, 0, sizeof(uint32_t)));
      }
    }
    if (av1t_ext_tx_ind[4][AV1B_DCT_DCT] != 3 || av1t_ext_tx_ind[5][AV1B_DCT_DCT] != 7) { set_error("transform-type symbol tables changed"); free_all(e); delete e; return AV1B_ERR_INTERNAL; }
  }
  for (int p = 0; p < 3; p++) {
    const size_t n = e->plane_elems[p] * F;
    A(cudaMalloc(&e->d_rec[p], n * 2));
    if (e->loop_filters) A(cudaMalloc(&e->d_deb[p], n * 2));
    A(cudaMalloc(&e->d_fin[p], (n + e->plane_elems[p]) * 2));
    if (err == cudaSuccess) {
      A(cudaMemset(e->d_rec[p], 0, n * 2));
      A(cudaMemset(e->d_fin[p], 0, (n + e->plane_elems[p]) * 2));
    }
  }
  A(cudaMalloc(&e->d_map_key, e->map_elems)); A(cudaMalloc(&e->d_map_inter, e->map_elems));
  A(cudaMalloc(&e->d_noise_hist, 4096 * sizeof(uint32_t))); A(cudaMallocHost(&e->h_noise_hist, 4096 * sizeof(uint32_t)));
  A(cudaEventCreate(&e->ev_noise));
  if (cfg->enable_qm) {
    A(cudaMalloc(&e->d_qm, sizeof(av1t_qm_sq)));
    A(cudaMemcpy(e->d_qm, av1t_qm_sq, sizeof(av1t_qm_sq), cudaMemcpyHostToDevice));
  }
  if (e->lr_on) A(cudaMalloc(&e->d_lr_sse, 3 * e->lr_n * F * sizeof(unsigned long long)));
  if (!e->intra_only) {
    for (int l = 0; l < 3; l++) {
      const size_t el = e->plane_elems[0] >> (2 * l);
      A(cudaMalloc(&e->d_pyr[l], el * (F + kSlotFrame0) * 2));
      if (err == cudaSuccess) A(cudaMemset(e->d_pyr[l], 0, el * (F + kSlotFrame0) * 2));
    }
    if (e->mctf_cfg) {
      for (int p = 0; p < 3; p++) {
        A(cudaMalloc(&e->d_hist_src[p], e->plane_elems[p] * kHist * 2));
        A(cudaMalloc(&e->d_flt[p], e->plane_elems[p] * F * 2));
        if (err == cudaSuccess) { A(cudaMemset(e->d_hist_src[p], 0, e->plane_elems[p] * kHist * 2)); A(cudaMemset(e->d_flt[p], 0, e->plane_elems[p] * F * 2)); }
      }
      A(cudaMalloc(&e->d_mvs_tf, e->map_elems * kMaxSearches * 4));
      A(cudaMalloc(&e->d_mv2_tf, (size_t)((coded_w + 31) / 32) * ((coded_h + 31) / 32) * kMaxSearches * 4));
    }
    const size_t n2 = (size_t)((coded_w + 31) / 32) * ((coded_h + 31) / 32);
    A(cudaMalloc(&e->d_mv2, n2 * F * 4));
    A(cudaMalloc(&e->d_mvs, e->map_elems * F * 4));
    {
      const size_t n1 = (size_t)((coded_w + 15) / 16) * ((coded_h + 15) / 16);
      A(cudaMalloc(&e->d_mv_tmp, n1 * kMaxSearches * 4));
      A(cudaMalloc(&e->d_hist, (size_t)kMaxSearches * 2049 * sizeof(uint32_t)));
    }
    if (err == cudaSuccess) A(cudaMemset(e->d_mvs, 0, e->map_elems * F * 4));
  }
  if (err == cudaSuccess) {
    // partition maps are constant: fixed square blocks with forced splits at the picture edge
    A(launch_partition_fixed(e->g, e->blk_log2, e->d_map_key, 1, e->stream));
    A(launch_partition_fixed(e->g, 4, e->d_map_inter, 1, e->stream));
    A(cudaStreamSynchronize(e->stream));
  }
  if (err != cudaSuccess) {
    set_error("device/pinned allocation failed: %s", cudaGetErrorString(err));
    free_all(e); delete e; return AV1B_ERR_NOMEM;
  }
  e->pool = new ThreadPool(e->host_threads);
  *out = e;
  return AV1B_OK;
}

void av1b_encoder_destroy(av1b_encoder* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device_id);
  cudaDeviceSynchronize();
  free_all(e);
  delete e;
}

// Batches of one call: the sources of batch k+1 are uploaded while the kernels of batch k run (their slot's
// device buffers were last read by batch k-1), and the host entropy-codes batch k-1 meanwhile.
// frames != nullptr: host sources; else order[] indexes the resident clip (device-to-device gather on the input stream).
// prev: the slot that holds the batch before this one in the same chunk (nullptr at a chunk start): its last picture is what
// the first picture's scene score is taken against.  The scores follow the upload on the input stream.
static int stage_any(av1b_encoder* e, Slot& s, const Slot* prev, const av1b_frame_src* frames, const uint32_t* order, uint32_t f0, int n) {
  if (frames) {
    const int rc = stage(e, s, frames + f0, n);
    if (rc != AV1B_OK) return rc;
  } else {
    s.has_score = false;
    CK(cudaEventRecord(s.ev_h2d, e->s_in));
    for (int b = 0; b < n; b++)
      for (int p = 0; p < 3; p++)
        CK(cudaMemcpyAsync(s.d_src[p] + (size_t)b * e->plane_elems[p], e->d_clip[p] + (size_t)order[f0 + b] * e->plane_elems[p],
                           e->plane_elems[p] * 2, cudaMemcpyDeviceToDevice, e->s_in));
    CK(cudaEventRecord(s.ev_src, e->s_in));
  }
  if (e->scene_on && !e->intra_only) {
    const uint16_t* last = (prev && prev->n_frames > 0) ? prev->d_src[0] + (size_t)(prev->n_frames - 1) * e->plane_elems[0] : nullptr;
    CK(launch_scene_score(e->g, s.d_src[0], e->plane_elems[0], last, n, s.d_score, e->s_in));
    CK(cudaMemcpyAsync(s.h_score, s.d_score, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->s_in));
    CK(cudaEventRecord(s.ev_score, e->s_in));
    e->kernel_launches += 1;
    s.has_score = true;
  }
  return AV1B_OK;
}

// Frame-level parameters of the three frame kinds for the structure in force
static void set_structure(av1b_encoder* e, int gop_period) {
  const av1b_config& c = e->cfg;
  e->gop_period = gop_period;
  e->base_q_idx = gop_period > 1 ? std::max(1, e->q_nominal - 8) : e->q_nominal;
  e->base_q_idx_nonref = std::min(255, e->q_nominal + 64);   // (+48 .. +96 are within 2 % of each other in BD-rate with the period-6 structure; +64 and +80 are best)
  e->mctf_on = e->mctf_cfg && gop_period > 1;
  av1b_select_frame_params(c.bit_depth, e->base_q_idx, AV1B_INTER_FRAME, e->loop_filters ? 1 : 0, &e->fp_inter);
  av1b_select_frame_params(c.bit_depth, e->base_q_idx_nonref, AV1B_INTER_FRAME, e->loop_filters ? 1 : 0, &e->fp_nonref);
  e->fp_nonref.non_reference = 1;
  set_qm_levels(e, &e->fp_inter); set_qm_levels(e, &e->fp_nonref);
  // no CDEF in the frames nobody predicts from: at their quantiser almost every block is skipped (CDEF leaves those
  // alone), the index per superblock costs more than the filter gains (-2 % bytes, -0.02 dB) and the kernel is saved
  e->fp_nonref.cdef_bits = 0;
  for (int i = 0; i < 8; i++) { e->fp_nonref.cdef_y_strength[i] = 0; e->fp_nonref.cdef_uv_strength[i] = 0; }
  for (Av1bFrameParams* f : {&e->fp_inter, &e->fp_nonref}) {
    if (e->lr_on) { f->lr_type[0] = AV1B_RESTORE_SWITCHABLE; f->lr_type[1] = f->lr_type[2] = AV1B_RESTORE_NONE; f->lr_unit_shift = 0; f->lr_uv_shift = 0; }
    f->tile_cols_log2 = e->g_inter.tile_cols_log2; f->tile_rows_log2 = e->g_inter.tile_rows_log2;
  }
}

// A closed chunk starts (the pipeline is empty): with config.gop_period == 0 the structure follows the noise level of the
// chunk's first picture -- where the quantiser is fine enough to code the noise (ac step < noise sum / 7.5) a flat
// P chain is the cheaper structure, otherwise the one-level hierarchy with temporally filtered anchors.
static int begin_chunk(av1b_encoder* e, Slot& first) {
  if (e->intra_only) return AV1B_OK;
  int gop = e->cfg.gop_period > 0 ? e->cfg.gop_period : kDefaultGopPeriod;
  if (e->gop_auto || e->cfg.film_grain > 0) {
    CK(cudaStreamWaitEvent(e->s_in, first.ev_src, 0));
    CK(launch_noise_hist(e->g, first.d_src[0], e->d_noise_hist, e->s_in));
    CK(cudaMemcpyAsync(e->h_noise_hist, e->d_noise_hist, 4096 * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->s_in));
    CK(cudaEventRecord(e->ev_noise, e->s_in));
    CK(cudaEventSynchronize(e->ev_noise));
    e->kernel_launches += 1;
    e->noise_b = av1b_noise_from_hist(e->h_noise_hist);
    const int acq = e->cfg.bit_depth == 8 ? av1t_ac_q_8[e->q_nominal] : av1t_ac_q_10[e->q_nominal];
    if (e->gop_auto && 2 * e->noise_b > 15 * acq) gop = 1;
  }
  // film grain synthesis: where the structure takes the noise out (temporally filtered anchors, skipped blocks elsewhere) the
  // decoder puts grain of three quarters of the measured strength back; sigma = 0.0010658 * noise_b samples at the source's
  // bit depth, a scaling value of 64 stands for sigma 1 in 8-bit units.  The P chain codes the noise itself: no grain there.
  e->grain_scaling = 0;
  if (e->cfg.film_grain > 0 && gop > 1)
    e->grain_scaling = (int)std::min<long long>(255, ((long long)e->noise_b * 51 + (500 << (e->cfg.bit_depth - 8))) / (1000 << (e->cfg.bit_depth - 8)));
  if (gop != e->gop_period) set_structure(e, gop);
  if (e->rc_on && e->token_path && (e->cdf_q[0] != e->base_q_idx || e->cdf_q[1] != e->base_q_idx_nonref)) {
    std::vector<uint8_t> img(tile_cdfs_size());
    tile_cdfs_default(e->base_q_idx, img.data());
    CK(cudaMemcpy(e->d_cdf_init, img.data(), img.size(), cudaMemcpyHostToDevice));
    tile_cdfs_default(e->base_q_idx_nonref, img.data());
    CK(cudaMemcpy(e->d_cdf_init_alt, img.data(), img.size(), cudaMemcpyHostToDevice));
    e->cdf_q[0] = e->base_q_idx; e->cdf_q[1] = e->base_q_idx_nonref;
  }
  return AV1B_OK;
}

// finish the batches in flight, oldest first, until at most `keep` remain
static int drain_to(av1b_encoder* e, int keep, av1b_packet_cb out_cb, av1b_progress_cb prog_cb, void* user, int64_t total_frames,
                    std::chrono::steady_clock::time_point t0) {
  const int S = e->n_slots;
  while (e->pipe_next - e->pipe_done > keep) {
    Slot& sl = e->slot[e->pipe_done % S];
    int rc = finish(e, sl, sl.staged, out_cb, prog_cb, user, total_frames, t0);
    if (rc != AV1B_OK) return rc;
    e->pipe_done++;
  }
  return AV1B_OK;
}

// drain = false (av1b_encode_stream): the call returns with up to n_slots - 1 batches still on the device; their packets
// come out during the next call or av1b_encode_flush, so consecutive parts of a chunk keep the pipeline full.
static int run_batches(av1b_encoder* e, const av1b_frame_src* frames, const uint32_t* order, uint32_t n_frames, int64_t first_index,
                       int64_t total_frames, av1b_packet_cb out_cb, av1b_progress_cb prog_cb, void* user, bool drain = true) {
  const auto t0 = std::chrono::steady_clock::now();
  const uint32_t B = (uint32_t)e->batch;
  const bool staged = frames != nullptr;
  const int S = e->n_slots, L = S - 1;
  int rc;
  {
    Slot& s0 = e->slot[e->pipe_next % S];
    if (e->pipe_next >= S) CK(cudaStreamWaitEvent(e->s_in, s0.ev_k1, 0));   // the batch that used this slot has read its sources
    const Slot* before = (e->chunk_pos > 0 && e->pipe_next > 0) ? &e->slot[(e->pipe_next - 1) % S] : nullptr;
    if ((rc = stage_any(e, s0, before, frames, order, 0, (int)std::min<uint32_t>(B, n_frames))) != AV1B_OK) return rc;
    if (e->chunk_pos == 0 && (rc = begin_chunk(e, s0)) != AV1B_OK) return rc;
  }
  Slot* last_staged = nullptr;
  for (uint32_t f0 = 0; f0 < n_frames; f0 += B) {
    const int nb = (int)std::min<uint32_t>(B, n_frames - f0);
    const int i = e->pipe_next;
    Slot& cur = e->slot[i % S];
    cur.staged = staged;
    if ((rc = launch(e, cur, cur, nb, first_index + f0)) != AV1B_OK) return rc;
    e->pipe_next++;
    last_staged = &cur;
    if (f0 + B < n_frames) {
      Slot& nx = e->slot[(i + 1) % S];
      if (i + 1 >= S) CK(cudaStreamWaitEvent(e->s_in, nx.ev_k1, 0));   // batch i+1-S has read that slot's sources
      if ((rc = stage_any(e, nx, &cur, frames, order, f0 + B, (int)std::min<uint32_t>(B, n_frames - f0 - B))) != AV1B_OK) return rc;
    }
    // at most L batches stay in flight: the slot the next batch is staged into belongs to a finished one
    if ((rc = drain_to(e, L, out_cb, prog_cb, user, total_frames, t0)) != AV1B_OK) return rc;
  }
  if (drain) {
    if ((rc = drain_to(e, 0, out_cb, prog_cb, user, total_frames, t0)) != AV1B_OK) return rc;
    for (Slot& sl : e->slot) collect_h2d(e, sl);
  } else if (staged && last_staged) {
    CK(cudaEventSynchronize(last_staged->ev_src));   // the caller's buffers have been read
  }
  return AV1B_OK;
}

int av1b_encode_chunk(av1b_encoder* e, const av1b_frame_src* frames, uint32_t n_frames, av1b_packet_cb out_cb,
                      av1b_progress_cb prog_cb, void* user) {
  if (!e || !frames || !out_cb) { set_error("null argument"); return AV1B_ERR_INVALID; }
  if (n_frames == 0) return AV1B_OK;
  CK(cudaSetDevice(e->cfg.device_id));
  reset_stats(e);
  { int rc0 = drain_to(e, 0, e->last_cb, nullptr, e->last_user, 0, std::chrono::steady_clock::now()); if (rc0 != AV1B_OK) return rc0; }
  e->chunk_pos = 0;               // a chunk is a closed GOP: it starts with a key frame
  return run_batches(e, frames, nullptr, n_frames, 0, n_frames, out_cb, prog_cb, user);
}

// Streaming variant: a chunk handed over in parts (bounded host memory for long chunks).  The part
// with first_part != 0 starts a new closed GOP (key frame); every call returns after its frames have
// been delivered to out_cb.
int av1b_encode_part(av1b_encoder* e, const av1b_frame_src* frames, uint32_t n_frames, int first_part,
                     int64_t first_frame_index, av1b_packet_cb out_cb, av1b_progress_cb prog_cb, void* user) {
  if (!e || !frames || !out_cb || n_frames == 0) { set_error("null argument"); return AV1B_ERR_INVALID; }
  CK(cudaSetDevice(e->cfg.device_id));
  { int rc0 = drain_to(e, 0, e->last_cb, nullptr, e->last_user, 0, std::chrono::steady_clock::now()); if (rc0 != AV1B_OK) return rc0; }
  if (first_part) { reset_stats(e); e->chunk_pos = 0; }
  return run_batches(e, frames, nullptr, n_frames, first_frame_index, 0, out_cb, prog_cb, user);
}

// Streaming without a drain between the parts of a chunk: the call returns once the frames have been uploaded; their
// packets may be delivered by a later av1b_encode_stream / av1b_encode_flush call (always in order).  A part with
// first_part != 0 first flushes what is in flight (to the callback of the call that submitted it).
int av1b_encode_stream(av1b_encoder* e, const av1b_frame_src* frames, uint32_t n_frames, int first_part,
                       int64_t first_frame_index, av1b_packet_cb out_cb, av1b_progress_cb prog_cb, void* user) {
  if (!e || !frames || !out_cb || n_frames == 0) { set_error("null argument"); return AV1B_ERR_INVALID; }
  CK(cudaSetDevice(e->cfg.device_id));
  if (first_part) {
    int rc = drain_to(e, 0, e->last_cb, nullptr, e->last_user, 0, std::chrono::steady_clock::now());
    if (rc != AV1B_OK) return rc;
    reset_stats(e); e->chunk_pos = 0;
  }
  e->last_cb = out_cb; e->last_user = user;
  return run_batches(e, frames, nullptr, n_frames, first_frame_index, 0, out_cb, prog_cb, user, false);
}

int av1b_encode_flush(av1b_encoder* e, av1b_packet_cb out_cb, av1b_progress_cb prog_cb, void* user) {
  if (!e) { set_error("null argument"); return AV1B_ERR_INVALID; }
  CK(cudaSetDevice(e->cfg.device_id));
  int rc = drain_to(e, 0, out_cb ? out_cb : e->last_cb, prog_cb, out_cb ? user : e->last_user, 0, std::chrono::steady_clock::now());
  for (Slot& sl : e->slot) collect_h2d(e, sl);
  return rc;
}

// ---- device-resident flow (bench: "inputs already resident in HBM") -----------------------------
int av1b_stage_frames(av1b_encoder* e, int slot, const av1b_frame_src* frames, uint32_t n_frames) {
  if (!e || !frames || slot < 0 || slot > 1 || n_frames == 0 || (int)n_frames > e->batch) { set_error("bad argument"); return AV1B_ERR_INVALID; }
  CK(cudaSetDevice(e->cfg.device_id));
  int rc = stage(e, e->slot[slot], frames, (int)n_frames);
  if (rc) return rc;
  e->slot[slot].n_frames = (int)n_frames;
  CK(cudaStreamSynchronize(e->s_in));
  return AV1B_OK;
}

int av1b_encode_resident(av1b_encoder* e, uint32_t n_steps, av1b_packet_cb out_cb, void* user) {
  if (!e || n_steps == 0) { set_error("bad argument"); return AV1B_ERR_INVALID; }
  if (e->slot[0].n_frames <= 0 || e->slot[1].n_frames <= 0) { set_error("stage both slots first"); return AV1B_ERR_INVALID; }
  CK(cudaSetDevice(e->cfg.device_id));
  reset_stats(e);
  e->chunk_pos = 0;
  const auto t0 = std::chrono::steady_clock::now();
  int rc;
  int64_t idx = 0;
  const int n0 = e->slot[0].n_frames, n1 = e->slot[1].n_frames;
  // sources alternate between the two staged slots; results go through all slots like in run_batches
  const int S = e->n_slots, L = S - 1;
  for (uint32_t i = 0; i < n_steps; i++) {
    const int nb = (i & 1) ? n1 : n0;
    if ((rc = launch(e, e->slot[i % S], e->slot[i & 1], nb, idx)) != AV1B_OK) return rc;
    idx += nb;
    if ((int)i >= L && (rc = finish(e, e->slot[(i - L) % S], false, out_cb, nullptr, user, 0, t0)) != AV1B_OK) return rc;
  }
  for (uint32_t k = n_steps > (uint32_t)L ? n_steps - L : 0; k < n_steps; k++)
    if ((rc = finish(e, e->slot[k % S], false, out_cb, nullptr, user, 0, t0)) != AV1B_OK) return rc;
  e->slot[0].n_frames = n0; e->slot[1].n_frames = n1;
  return AV1B_OK;
}

// ---- resident clip: n source pictures uploaded once, then closed chunks are coded out of them by index ----
int av1b_stage_clip(av1b_encoder* e, const av1b_frame_src* frames, uint32_t n) {
  if (!e || !frames || n == 0) { set_error("bad argument"); return AV1B_ERR_INVALID; }
  if (e->padded_src) { set_error("the resident clip takes sizes that are multiples of 8"); return AV1B_ERR_INVALID; }
  CK(cudaSetDevice(e->cfg.device_id));
  CK(cudaDeviceSynchronize());
  if (n != e->n_clip) {
    for (int p = 0; p < 3; p++) { cudaFree(e->d_clip[p]); e->d_clip[p] = nullptr; }
    e->n_clip = 0;
    for (int p = 0; p < 3; p++)
      if (cudaMalloc(&e->d_clip[p], e->plane_elems[p] * n * 2) != cudaSuccess) { cudaGetLastError(); set_error("resident clip: out of device memory"); return AV1B_ERR_NOMEM; }
    e->n_clip = n;
  }
  for (int p = 0; p < 3; p++) CK(cudaMemset(e->d_clip[p], 0, e->plane_elems[p] * n * 2));
  const Av1bGeom& g = e->g;
  for (uint32_t b = 0; b < n; b++)
    for (int p = 0; p < 3; p++) {
      const int w = p ? g.width >> 1 : g.width, h = p ? g.height >> 1 : g.height;
      CK(cudaMemcpy2D(e->d_clip[p] + (size_t)b * e->plane_elems[p], (size_t)g.stride[p] * 2, frames[b].planes[p],
                      (size_t)frames[b].stride[p] * 2, (size_t)w * 2, h, cudaMemcpyHostToDevice));
    }
  return AV1B_OK;
}

int av1b_encode_clip(av1b_encoder* e, const uint32_t* order, uint32_t n_frames, int accumulate_stats, av1b_packet_cb out_cb, void* user) {
  if (!e || !order || n_frames == 0) { set_error("bad argument"); return AV1B_ERR_INVALID; }
  if (!e->n_clip) { set_error("stage a clip first (av1b_stage_clip)"); return AV1B_ERR_INVALID; }
  for (uint32_t i = 0; i < n_frames; i++) if (order[i] >= e->n_clip) { set_error("clip index out of range"); return AV1B_ERR_INVALID; }
  CK(cudaSetDevice(e->cfg.device_id));
  { int rc0 = drain_to(e, 0, e->last_cb, nullptr, e->last_user, 0, std::chrono::steady_clock::now()); if (rc0 != AV1B_OK) return rc0; }
  if (!accumulate_stats) reset_stats(e);
  e->chunk_pos = 0;               // a closed chunk: it starts with a key frame
  return run_batches(e, nullptr, order, n_frames, 0, n_frames, out_cb, nullptr, user);
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

int av1b_get_frame_syms(av1b_encoder* e, uint32_t frame, Av1bBlockInfo* blocks, int16_t* const coef[3]) {
  if (!e || !e->keep || frame >= e->kept.size()) { set_error("symbols not kept or bad index"); return AV1B_ERR_INVALID; }
  if (blocks) memcpy(blocks, e->kept[frame].blocks.data(), e->map_elems * sizeof(Av1bBlockInfo));
  if (coef) for (int p = 0; p < 3; p++) if (coef[p]) memcpy(coef[p], e->kept[frame].coef[p].data(), e->plane_elems[p] * 2);
  return AV1B_OK;
}

int av1b_get_frame_params(av1b_encoder* e, Av1bFrameParams* fp) {
  if (!e || !fp) return AV1B_ERR_INVALID;
  *fp = e->fp_key;
  return AV1B_OK;
}

int av1b_get_inter_frame_params(av1b_encoder* e, Av1bFrameParams* fp) {
  if (!e || !fp) return AV1B_ERR_INVALID;
  *fp = e->fp_inter;
  return AV1B_OK;
}

int av1b_get_class_params(av1b_encoder* e, int kind, Av1bFrameParams* fp) {
  if (!e || !fp || kind < 0 || kind > 2) return AV1B_ERR_INVALID;
  *fp = kind_params(e, kind);
  return AV1B_OK;
}

int av1b_get_quality(av1b_encoder* e, double* psnr_y, double* ssim_y, int64_t* frames) {
  if (!e) return AV1B_ERR_INVALID;
  if (!e->quality_on) { set_error("quality statistics are off (config.tune[3] = 1 switches them on)"); return AV1B_ERR_INVALID; }
  if (psnr_y) *psnr_y = e->q_frames > 0 ? e->q_psnr_sum / e->q_frames : 0.0;
  if (ssim_y) *ssim_y = e->q_frames > 0 ? e->q_ssim / e->q_frames : 0.0;
  if (frames) *frames = (int64_t)e->q_frames;
  return AV1B_OK;
}

int av1b_get_chunk_info(av1b_encoder* e, int32_t info[8]) {
  if (!e || !info) return AV1B_ERR_INVALID;
  const int32_t v[8] = {e->gop_period, e->base_q_idx_key, e->base_q_idx, e->base_q_idx_nonref, e->mctf_on ? 1 : 0, e->noise_b, e->q_nominal, e->gop_auto ? 1 : 0};
  for (int i = 0; i < 8; i++) info[i] = v[i];
  return AV1B_OK;
}

int av1b_get_frame_kind(av1b_encoder* e, int64_t pos_in_chunk) {
  if (!e || pos_in_chunk < 0) return AV1B_ERR_INVALID;
  // debug mode knows what the frame was coded as (a scene change inside the chunk restarts the structure)
  if (e->keep && (size_t)pos_in_chunk < e->kept.size()) return e->kept[(size_t)pos_in_chunk].kind;
  return frame_kind(e, pos_in_chunk);
}

int av1b_get_me_lambda(av1b_encoder* e) {
  if (!e) return AV1B_ERR_INVALID;
  return (e->cfg.bit_depth == 8 ? av1t_ac_q_8[e->base_q_idx] : av1t_ac_q_10[e->base_q_idx]) >> 1;
}

int av1b_get_frame_is_key(av1b_encoder* e, uint32_t frame) {
  if (!e || !e->keep || frame >= e->kept.size()) return AV1B_ERR_INVALID;
  return e->kept[frame].is_key;
}

int av1b_get_cdef_idx(av1b_encoder* e, uint32_t frame, uint8_t* idx) {
  if (!e || !idx || !e->keep || frame >= e->kept.size()) { set_error("cdef_idx not kept or bad index"); return AV1B_ERR_INVALID; }
  memcpy(idx, e->kept[frame].cdef_idx.data(), e->kept[frame].cdef_idx.size());
  return AV1B_OK;
}

int av1b_get_lr_units(av1b_encoder* e, uint32_t frame, Av1bLrUnit* units, int32_t* rows, int32_t* cols) {
  if (!e || !e->lr_on || !e->keep || frame >= e->kept.size()) { set_error("restoration units not kept (preset <= 5 and config.reserved[0] = 1) or bad index"); return AV1B_ERR_INVALID; }
  if (rows) *rows = e->lr_rows;
  if (cols) *cols = e->lr_cols;
  if (units) memcpy(units, e->kept[frame].lr_units.data(), e->lr_n * sizeof(Av1bLrUnit));
  return AV1B_OK;
}

int av1b_get_geom(av1b_encoder* e, Av1bGeom* g) {
  if (!e || !g) return AV1B_ERR_INVALID;
  *g = e->g;
  return AV1B_OK;
}

void* av1b_host_alloc(int device, size_t bytes) {
  void* p = nullptr;
  if (bytes == 0) return nullptr;
  if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", device); cudaGetLastError(); return nullptr; }
  const cudaError_t err = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
  if (err != cudaSuccess) { set_error("cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(err)); cudaGetLastError(); return nullptr; }
  return p;
}

void av1b_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int av1b_get_stats(av1b_encoder* e, double* stats, int n) {
  if (!e || !stats) return AV1B_ERR_INVALID;
  const double v[24] = {e->t_h2d_ms, e->t_kernel_ms, e->t_d2h_ms, e->t_pack_ms, (double)e->kernel_launches,
                        (double)e->base_q_idx, e->t_intra_ms, (double)e->intra_launches, (double)e->frames_done,
                        (double)e->bytes_out, e->t_deblock_ms, e->t_cdef_ms, e->t_inter_ms, e->t_me_ms,
                        (double)e->inter_launches, (double)e->key_frames, (double)e->staged_direct, e->t_tok_ms, (double)e->n_tokens, (double)e->d2h_bytes, e->t_lr_ms, e->t_rc_ms, e->t_mctf_ms, (double)e->mctf_frames};
  for (int i = 0; i < n && i < 24; i++) stats[i] = v[i];
  return AV1B_OK;
}

}  // extern "C"
