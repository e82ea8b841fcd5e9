/* av1b200 -- C ABI of the B200-native AV1 encode backend.
 *
 * This is the boundary a reference-side FFI binds (BASELINE.json north_star: "a new Rust crate
 * (av1-cuda-sys plus a safe wrapper) calls hand-written CUDA kernels through a thin C-ABI").
 * It replaces what the reference reaches by exec'ing `av1an`:
 *   /root/reference/crates/daemon/src/encode/av1an.rs:79-107  build_av1an_command  -> av1b_config
 *   /root/reference/crates/daemon/src/encode/av1an.rs:126-139 run_av1an            -> av1b_encode_chunk
 *   /root/reference/crates/daemon/src/encode/av1an.rs:18-30   EncodeError          -> negative return codes
 *   /root/reference/crates/daemon/src/startup.rs:98-116       `av1an --version`    -> av1b_version
 * Conventions: 0 = OK, negative = error (av1b_last_error() gives the thread-local message); no
 * exceptions or aborts cross the boundary; plain pointers and sizes only.  The caller owns input
 * buffers for the duration of a call; packet memory handed to a callback is valid until it returns.
 * One encoder handle = one GPU + one stream set; handles are independent and may be driven from
 * different threads (one chunk stream per GPU).
 */
#ifndef AV1B200_H_
#define AV1B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AV1B_OK 0
#define AV1B_ERR_INVALID (-1)      /* bad argument / unsupported configuration           */
#define AV1B_ERR_NO_DEVICE (-2)    /* no CUDA device: there is NO CPU fallback           */
#define AV1B_ERR_CUDA (-3)         /* CUDA runtime error                                  */
#define AV1B_ERR_NOMEM (-4)
#define AV1B_ERR_CALLBACK (-5)     /* packet callback returned non-zero                   */
#define AV1B_ERR_INTERNAL (-6)

/* Mirrors the knobs the daemon passes through `--video-params` (av1an.rs:14 SVT_PARAMS) plus geometry. */
typedef struct av1b_config {
  int32_t width, height;          /* luma samples of the source, 16..8192 x 16..4352.  Sizes that are not multiples of 8 are coded
                                     padded (edge replication) to the next multiple; the frame headers then carry the source size
                                     as render_size and decoders output the padded frame (av1b_get_recon / av1b_get_geom too)   */
  int32_t bit_depth;              /* 8 or 10 (input samples are uint16 either way)                 */
  int32_t fps_num, fps_den;
  int32_t crf;                    /* 0..63, SVT-AV1 --crf                                          */
  int32_t preset;                 /* SVT-AV1 --preset: <= 5 adds loop restoration (per-unit decision), <= 3 a third sweep of the
                                     superblock-level regularisation of the vector field (-4 % bytes at equal PSNR) */
  int32_t keyint;                 /* --keyint                                                      */
  int32_t lookahead;              /* --lookahead: source pictures the temporal filter of key / anchor pictures may look ahead
                                     (it uses up to 4); -1 = default, 0 = none                                        */
  int32_t film_grain;             /* --film-grain 0..50: strength of the temporal (denoising) filter on top of what the
                                     quantiser asks for; > 0 also signals film grain parameters (spec 5.9.30: flat luma scaling
                                     from the noise level measured on the chunk's first picture, chroma from luma, white grain)
                                     in the chunks that are coded with the filtered structure, so that decoders put back
                                     what the filter and the skipped blocks took out                                  */
  int32_t enable_qm, qm_min, qm_max; /* --enable-qm 1 --qm-min A --qm-max B (0 <= A <= B <= 15): quantisation matrices (spec 7.12.3) at the
                                     level the frame's quantiser index maps to, A + qindex * (B + 1 - A) / 256 as in SVT-AV1 / libaom
                                     (0 steepest .. 15 flat), luma and chroma alike; defaults 0 / 8 / 15                */
  int32_t tile_cols_log2, tile_rows_log2; /* -1 = auto (fill the GPU)                              */
  int32_t device_id;
  int32_t hdr;                    /* 1: signal BT.2020/PQ in the sequence header                   */
  int32_t host_threads;           /* entropy-coding threads, 0 = auto                              */
  int32_t frames_in_flight;       /* frames batched per device pass, 0 = auto                      */
  int32_t gop_period;             /* one-level hierarchy: every gop_period-th frame after a key frame is an anchor (inter frame that
                                     becomes the reference, quantiser index - 8); the frames between two anchors predict from the last
                                     anchor at quantiser index + 64 and are referenced by nobody.  1 = plain P chain.
                                     0 = chosen per chunk: 6, or the P chain where the quantiser is fine enough to code the source's noise
                                     (noise estimate of the chunk's first picture against the quantiser step) */
  int32_t tune[7];                /* [0]: 1 = vector-field regularisation of the motion search off; [1]: 1 = fixed 16x16 key-frame
                                     partition (default: 64x64 / 32x32 blocks where the source is smooth); [2]: 1 = temporal filter of
                                     key / anchor source pictures off; [3]: 1 = PSNR / SSIM of every frame on the device (av1b_get_quality);
                                     [4]: 1 = no key frame at scene changes inside a chunk (default: scene scores are computed on the GPU as the
                                     pictures arrive and a change restarts the structure) */
  int32_t reserved[8];            /* [0]: keep recon+symbols per frame (tests); [1]: fixed block log2 (3..6), 0 = default;
                                     [2]: 1 = in-loop filters off; [3]: 1 = every frame is a key frame;
                                     [4]: inter transform-block drop threshold (0 = off);
                                     [5]: inter-frame entropy coding path: 0 = device tokenizer, range coder on the device when this encoder has fewer
                                          than 12 host threads, else on the host (default); 3 = device tokenizer + host range coder; 4 = device
                                          tokenizer + device range coder;
                                          1 = host block walker over raster levels, 2 = host block walker over in-place packed symbols;
                                     [6]: 1 = loop restoration off whatever the preset;
                                     [7]: target inter-frame tile size in superblocks (0 = default: 6 with the device range coder, else 12) */
} av1b_config;

typedef struct av1b_encoder av1b_encoder;

/* One frame of 4:2:0 source, samples as uint16 (8-bit content in the low byte). Strides in samples. */
typedef struct av1b_frame_src {
  const uint16_t* planes[3];
  int32_t stride[3];
} av1b_frame_src;

/* Called once per temporal unit (low-overhead OBU format), in display order. Non-zero aborts. */
typedef int (*av1b_packet_cb)(void* user, const uint8_t* data, size_t size, int64_t frame_index, int is_key);
/* Progress: frames done so far of the chunk (feeds JobMetrics.frames_encoded / fps, metrics.rs:16-24). */
typedef void (*av1b_progress_cb)(void* user, int64_t frames_done, int64_t frames_total, double fps);

int av1b_version(char* buf, size_t cap);
int av1b_device_count(void);
const char* av1b_last_error(void);

void av1b_config_default(av1b_config* cfg);
int av1b_encoder_create(const av1b_config* cfg, av1b_encoder** out);
void av1b_encoder_destroy(av1b_encoder* enc);
/* Encodes n_frames as one closed chunk: first TU carries the sequence header + key frame. */
int av1b_encode_chunk(av1b_encoder* enc, const av1b_frame_src* frames, uint32_t n_frames,
                      av1b_packet_cb out_cb, av1b_progress_cb prog_cb, void* user);
/* Streaming form of av1b_encode_chunk for long chunks: the chunk is handed over in parts; the part
 * with first_part != 0 starts a new closed GOP.  frame_index passed to out_cb = first_frame_index + k. */
int av1b_encode_part(av1b_encoder* enc, const av1b_frame_src* frames, uint32_t n_frames, int first_part,
                     int64_t first_frame_index, av1b_packet_cb out_cb, av1b_progress_cb prog_cb, void* user);
/* Like av1b_encode_part, but without draining the pipeline between the parts of a chunk: the call returns as soon as its
 * frames have been uploaded (the caller may reuse the buffers); packets are delivered in order by this or a later
 * av1b_encode_stream / av1b_encode_flush call, to the callback of the call that delivers them (a part with first_part != 0
 * first delivers everything still in flight to the callback of the call that submitted it). */
int av1b_encode_stream(av1b_encoder* enc, const av1b_frame_src* frames, uint32_t n_frames, int first_part,
                       int64_t first_frame_index, av1b_packet_cb out_cb, av1b_progress_cb prog_cb, void* user);
/* Delivers every packet still in flight (out_cb NULL: to the callback of the last av1b_encode_stream call). */
int av1b_encode_flush(av1b_encoder* enc, av1b_packet_cb out_cb, av1b_progress_cb prog_cb, void* user);
/* Reconstruction of the most recently encoded frame `frame_in_chunk` of the last chunk (post loop
 * filter), for the recon-vs-decode check. dst planes: uint16, strides in samples. */
int av1b_get_recon(av1b_encoder* enc, uint32_t frame_in_chunk, uint16_t* const dst[3], const int32_t stride[3]);

/* ---- device-resident flow: upload up to frames_in_flight frames into slot 0 and slot 1 once, then
 * encode them n_steps times (slot = step & 1) with the same pipelining as av1b_encode_chunk. Used to
 * measure throughput with the inputs already resident in HBM. out_cb may be NULL. ---- */
int av1b_stage_frames(av1b_encoder* enc, int slot, const av1b_frame_src* frames, uint32_t n_frames);
int av1b_encode_resident(av1b_encoder* enc, uint32_t n_steps, av1b_packet_cb out_cb, void* user);

/* ---- resident clip: upload n source pictures into HBM once (av1b_stage_clip), then code closed chunks made of them:
 * chunk frame i is clip picture order[i] (device-to-device gather into the pipeline's batch buffers, same pipelining as
 * av1b_encode_chunk).  accumulate_stats != 0 keeps adding to the statistics of the previous call.  out_cb may be NULL. ---- */
int av1b_stage_clip(av1b_encoder* enc, const av1b_frame_src* frames, uint32_t n_frames);
int av1b_encode_clip(av1b_encoder* enc, const uint32_t* order, uint32_t n_frames, int accumulate_stats, av1b_packet_cb out_cb, void* user);

/* ---- diagnostics (need config.reserved[0] = 1: keep per-frame reconstruction and symbols) ---- */
struct Av1bBlockInfo; struct Av1bGeom;
int av1b_get_frame_syms(av1b_encoder* enc, uint32_t frame_in_chunk, struct Av1bBlockInfo* blocks, int16_t* const coef[3]);
int av1b_get_geom(av1b_encoder* enc, struct Av1bGeom* geom);
struct Av1bFrameParams;
/* frame-level parameters the encoder signals for key frames (deblock levels, CDEF presets, ...) */
int av1b_get_frame_params(av1b_encoder* enc, struct Av1bFrameParams* fp);
int av1b_get_inter_frame_params(av1b_encoder* enc, struct Av1bFrameParams* fp);
/* config.tune[3] = 1: mean luma PSNR (dB) and SSIM (non-overlapping 8x8 windows) of the frames coded since the last chunk
 * start, reconstruction against the source as handed in (JobMetrics.psnr / .ssim, metrics.rs:12-30) */
int av1b_get_quality(av1b_encoder* enc, double* psnr_y, double* ssim_y, int64_t* frames);
/* structure of the chunk coded last: info[0..7] = gop_period in force, quantiser index of key / anchor / non-reference frames,
 * temporal filter on, noise estimate of the chunk's first picture, quantiser index of the CRF, structure chosen automatically */
int av1b_get_chunk_info(av1b_encoder* enc, int32_t info[8]);
/* noise level from the 4096-bin histogram of 16x16-block noise sums (lower quartile outside bin 0); pure function */
int av1b_noise_from_hist(const uint32_t* hist4096);
/* kind of the frame at position pos of a closed GOP: 0 key, 1 anchor, 2 non-reference; and the frame-level parameters of a kind */
int av1b_get_frame_kind(av1b_encoder* enc, int64_t pos_in_chunk);
int av1b_get_class_params(av1b_encoder* enc, int kind, struct Av1bFrameParams* fp);
/* vector-deviation cost (SAD units) the encoder uses in its motion search: half the AC quantiser step */
int av1b_get_me_lambda(av1b_encoder* enc);
/* 1 / 0: whether kept frame `frame_in_chunk` was coded as a key frame (needs config.reserved[0] = 1) */
int av1b_get_frame_is_key(av1b_encoder* enc, uint32_t frame_in_chunk);
/* luma restoration units of a kept frame (preset <= 5: loop restoration on): units[rows*cols] (may be NULL to query the grid) */
struct Av1bLrUnit;
int av1b_get_lr_units(av1b_encoder* enc, uint32_t frame_in_chunk, struct Av1bLrUnit* units, int32_t* rows, int32_t* cols);
/* chosen CDEF preset per 64x64 superblock of a kept frame: idx[sb_rows*sb_cols] */
int av1b_get_cdef_idx(av1b_encoder* enc, uint32_t frame_in_chunk, uint8_t* idx);
/* pure function (no device): deblock levels and CDEF presets from bit depth / quantiser / frame type */
int av1b_select_frame_params(int bit_depth, int base_q_idx, int frame_type, int loop_filters, struct Av1bFrameParams* fp);
/* Page-locked host memory for source frames.  Planes that lie in page-locked memory (from here, or the
 * caller's own cudaHostAlloc / cudaHostRegister) are read by the copy engine where they are; pageable
 * planes go through one extra host copy into the encoder's own staging buffer.  The allocation is made in
 * the context of `device` (one the caller encodes on) and is usable from every device.  NULL on failure. */
void* av1b_host_alloc(int device, size_t bytes);
void av1b_host_free(void* p);
/* stats[0..23] = h2d_ms, kernel_ms, d2h_ms, pack_ms, kernel_launches, base_q_idx, intra_kernel_ms,
 * intra_kernel_launches, frames_done, bytes_out, deblock_ms, cdef_ms, inter_kernel_ms, me_ms (pyramid + search),
 * inter_kernel_launches, key_frames, frames uploaded straight from page-locked caller memory, tokenizer_ms,
 * tokens produced, bytes copied device -> host, loop_restoration_ms, range_coder_ms (overlaps the next batch), temporal_filter_ms (its searches + filter), pictures filtered, of the last chunk / resident run (CUDA-event times) */
int av1b_get_stats(av1b_encoder* enc, double* stats, int n);

/* ---- kernel suite (BASELINE.json config 2 "kernel bit-exact suite") -----------------------------
 * Each call uploads the host buffers to `device`, runs ONE CUDA kernel `reps` times between CUDA
 * events (mean milliseconds per launch -> *ms_per_launch, may be NULL) and downloads the result.
 * Frame planes use the padded layout of Av1bGeom (rows[p] x stride[p] uint16 samples per plane,
 * n_frames planes back to back); `blocks` is [n_frames][h8*w8].
 * These are the units libaom's C reference (av1_inv_txfm2d_add_*_c, aom_highbd_lpf_*_c, cdef_*_c,
 * restoration) is compared with; they replace arithmetic behind av1an.rs:126-139. */
struct Av1bLrUnit;
/* coef: n_blocks x 1024 int32 (row-major, min(w,32) values per row); dst: n_blocks x (w*h) prediction in,
 * reconstruction out */
int av1b_k_inv_txfm_add(int device, const int32_t* coef, uint16_t* dst, int n_blocks, int w, int h, int tx_type,
                        int bit_depth, int reps, double* ms_per_launch);
/* Encoder-side forward transform (E4), every size (4x4 .. 64x64, 2:1 and 4:1 rectangles) and every legal type
 * (DCT / ADST / flipADST / identity): block b reads resid + b*w*h (row-major int16 residual) and writes
 * coef + b * min(w,32) * min(h,32) (row-major, spec layout; what dequantisation + av1b_k_inv_txfm_add invert). */
int av1b_k_fwd_txfm(int device, const int16_t* resid, int32_t* coef, int n_blocks, int w, int h, int tx_type, int reps,
                    double* ms_per_launch);
int av1b_k_deblock(int device, int width, int height, int bit_depth, int n_frames, const struct Av1bBlockInfo* blocks,
                   const uint16_t* const in[3], uint16_t* const out[3], const int32_t lf_level[4], int sharpness,
                   int reps, double* ms_per_launch);
/* src != NULL: decide the preset per superblock against the source (written to cdef_idx_out);
 * forced_idx != NULL: apply the given presets (normative filter only). */
int av1b_k_cdef(int device, int width, int height, int bit_depth, int n_frames, const struct Av1bBlockInfo* blocks,
                const struct Av1bFrameParams* fp, const uint16_t* const in[3], const uint16_t* const src[3],
                const uint8_t* forced_idx, uint16_t* const out[3], uint8_t* cdef_idx_out, int reps,
                double* ms_per_launch);
int av1b_k_lr(int device, int width, int height, int bit_depth, int n_frames, const struct Av1bFrameParams* fp,
              const uint16_t* const cdef[3], const uint16_t* const deb[3], const struct Av1bLrUnit* const units[3],
              uint16_t* const out[3], int reps, double* ms_per_launch);

/* Encoder-side loop-restoration decision (E8 "search"), luma plane, 64x64 units, frame type SWITCHABLE: every unit
 * picks NONE, WIENER with cand's taps or SGRPROJ with cand's set / weights, whichever has the smallest squared error
 * against the source; a restoring candidate must beat NONE by more than `bias`.  units_out: [rows*cols];
 * sse_out (may be NULL): [3][rows*cols] (none, Wiener, self-guided). */
int av1b_k_lr_search(int device, int width, int height, int bit_depth, const struct Av1bLrUnit* cand,
                     const uint16_t* const cdef[3], const uint16_t* const deb[3], const uint16_t* src_y, int64_t bias,
                     struct Av1bLrUnit* units_out, uint64_t* sse_out, int reps, double* ms_per_launch);

/* Source pyramid: l0 = n_frames padded luma planes; l1 / l2 = the 1/2 and 1/4 planes
 * (strides stride0/2, stride0/4). */
int av1b_k_pyramid(int device, int width, int height, int n_frames, const uint16_t* l0, uint16_t* l1, uint16_t* l2,
                   int reps, double* ms_per_launch);
/* Hierarchical motion estimation (E2): cur / ref = n_frames padded luma planes each; mv_out =
 * [n_frames][h8*w8][2] (row, col) in 1/8 luma samples. lambda = cost of one sample of deviation from the
 * parent vector (SAD units). The quarter-resolution level compares 8-bit samples, min(v >> (bit_depth - 8), 255).
 * The timed launch is the search (two kernels). */
int av1b_k_hme(int device, int width, int height, int bit_depth, int n_frames, const uint16_t* cur_l0, const uint16_t* ref_l0,
               int lambda, int16_t* mv_out, int reps, double* ms_per_launch);
/* av1b_k_hme followed by `passes` checkerboard sweeps of the superblock-level rate-distortion regularisation of the
 * vector field (E2): per 64x64 superblock the SADs (bilinear quarter-sample interpolation) of its 16x16 blocks at every
 * candidate vector (zero, the frame's dominant vector, the neighbouring superblocks', its own), a relaxation of the blocks
 * with lam_s per neighbour that holds another vector, then one vector for the whole superblock / a 32x32 quadrant where
 * that is cheaper at lam_r per bit of vector rate. */
int av1b_k_hme_sbrd(int device, int width, int height, int bit_depth, int n_frames, const uint16_t* cur_l0, const uint16_t* ref_l0,
                    int lambda, int lam_s, int lam_r, int passes, int16_t* mv_out, int reps, double* ms_per_launch);
/* Motion-compensated temporal filter of one source picture (encoder side; `--film-grain` denoising and `--lookahead`,
 * av1an.rs:14): weighted mean of the picture and n_nb <= 6 neighbours in time, neighbour k compensated onto the picture
 * with the normative interpolation and the vectors mvs[k] ([h8*w8][2], from av1b_k_hme of the picture against it).
 * nb_planes: [n_nb * 3] padded planes.  thr_b / thr_p: the block weight falls to zero at this luma mean squared
 * error, the sample weight at this squared difference. */
int av1b_k_mctf(int device, int width, int height, int bit_depth, const uint16_t* const cur[3], int n_nb,
                const uint16_t* const* nb_planes, const int16_t* const* mvs, int thr_b, int thr_p, uint16_t* const out[3],
                int reps, double* ms_per_launch);
/* Noise level of one padded luma plane (structure decision, temporal filter): lower-quartile 16x16-block sum of the
 * noise-mask response |I * [1 -2 1; -2 4 -2; 1 -2 1]| over the 14x14 inner samples; sigma ~= 0.0010658 * result. */
int av1b_k_noise_estimate(int device, int width, int height, const uint16_t* src_y, int32_t* noise_out);
/* Key-frame partition by smoothness (E3/E4 decision): src_y = one padded luma plane; map_out[h8*w8] = block log2 (3..6)
 * per 8x8 unit: 64x64 / 32x32 where the 4x4 box sums stay within thr of a plane, else 16x16 (8x8 at the picture edge). */
int av1b_k_partition_smooth(int device, int width, int height, const uint16_t* src_y, int thr, uint8_t* map_out);
/* Inter frame encode (E4 + E5 for inter frames): one frame; part_map [h8*w8] with values 3 / 4.
 * tb_zero_thr: drop transform blocks whose levels sum to <= thr; merge_skip: afterwards merge skipped
 * siblings with equal vectors into 32x32 / 64x64 blocks. */
int av1b_k_inter_encode(int device, int width, int height, int bit_depth, int base_q_idx, const uint8_t* part_map,
                        const int16_t* mvs, const uint16_t* const src[3], const uint16_t* const ref[3],
                        uint16_t* const rec[3], int16_t* const coef[3], struct Av1bBlockInfo* blocks, int tb_zero_thr,
                        int merge_skip, int reps, double* ms_per_launch);

/* ---- host entropy coder over symbol streams (the part that "runs on the host") -------------- */
struct Av1bSeqParams; struct Av1bFrameParams; struct Av1bFrameSyms;
int av1b_pack_sequence_header(const struct Av1bSeqParams* seq, uint8_t* out, size_t cap, size_t* len);
int av1b_pack_frame(const struct Av1bSeqParams* seq, const struct Av1bFrameParams* fp,
                    const struct Av1bFrameSyms* syms, int n_threads, int with_temporal_delimiter,
                    uint8_t* out, size_t cap, size_t* len);

/* The same inter frame through the token path ("device-produced symbol streams"): the CPU
 * statement of the device tokenizer derives one token per coded symbol, the tokens are range-coded tile by
 * tile.  Gives the bytes of av1b_pack_frame; n_tokens (may be NULL) = tokens of the frame. */
int av1b_pack_frame_tokens(const struct Av1bSeqParams* seq, const struct Av1bFrameParams* fp,
                           const struct Av1bFrameSyms* syms, int with_temporal_delimiter,
                           uint8_t* out, size_t cap, size_t* len, uint64_t* n_tokens);

#ifdef __cplusplus
}
#endif
#endif /* AV1B200_H_ */
