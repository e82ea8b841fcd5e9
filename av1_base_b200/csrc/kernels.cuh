// Device-side interface of the encode kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "av1b_types.h"

namespace av1b {

// Per-launch parameters of the closed-loop intra encode kernel (one CTA per tile per frame).
struct IntraLaunch {
  Av1bGeom g;
  int32_t bit_depth;
  int32_t base_q_idx;
  int32_t quant_rnd;
  int32_t dc_q, ac_q;
  // frame-batch layout: plane p of frame f starts at base[p] + f * frame_stride[p] (in elements)
  const uint16_t* src[3];
  uint16_t* rec[3];
  int16_t* coef[3];
  size_t plane_elems[3];
  Av1bBlockInfo* blocks;      // [n_frames][h8*w8]
  const uint8_t* part_map;    // [n_frames][h8*w8]   (per-frame partition quadtree)
  size_t map_elems;
  // quantisation matrices (spec 7.12.3): luma / chroma weights of the frame's level (av1_qm_tables.h: the four square
  // sizes, 1360 bytes, device memory) or nullptr = flat
  const uint8_t* qm[2] = {nullptr, nullptr};
};

// Deblocking filter (deblock_kernel.cu): in -> out, one CTA per superblock.
struct DeblockLaunch {
  Av1bGeom g;
  int32_t bit_depth;
  int32_t lf_level[4];        // Y vertical edges, Y horizontal edges, U, V
  int32_t sharpness;
  const uint16_t* in[3];
  uint16_t* out[3];
  size_t plane_elems[3];
  const Av1bBlockInfo* blocks;
  size_t map_elems;
};

// CDEF with per-superblock preset decision (cdef_kernel.cu): in = deblocked frame.
struct CdefLaunch {
  Av1bGeom g;
  int32_t bit_depth;
  int32_t cdef_damping, cdef_bits;
  int32_t y_strength[8], uv_strength[8];
  const uint16_t* in[3];
  const uint16_t* src[3];     // source frame (decision only; may alias `in` when cdef_bits == 0)
  uint16_t* out[3];
  size_t plane_elems[3];
  const Av1bBlockInfo* blocks;
  size_t map_elems;
  uint8_t* cdef_idx;          // [n_frames][sb_rows*sb_cols] (chosen preset) or nullptr
  const uint8_t* forced_idx;  // non-null: use these presets instead of deciding (kernel suite)
};

// Loop restoration (lr_kernel.cu): cdef = CDEF output, deb = deblocked (pre-CDEF) frame.
struct LrLaunch {
  Av1bGeom g;
  int32_t bit_depth;
  int32_t lr_type[3];
  int32_t unit_size[3], unit_rows[3], unit_cols[3];
  const uint16_t* cdef[3];
  const uint16_t* deb[3];
  uint16_t* out[3];
  size_t plane_elems[3];
  const Av1bLrUnit* units[3]; // [n_frames][unit_rows*unit_cols] per plane, or nullptr
  // decision (launch_lr_search), luma only:
  const uint16_t* src_y;      // source luma
  Av1bLrUnit cand;            // the Wiener taps and the self-guided set / weights every unit may choose
  unsigned long long* sse;    // scratch [n_frames][3][unit_rows*unit_cols]: squared errors of none / Wiener / self-guided
};

// Batched normative inverse transform + reconstruction, all 19 AV1 sizes (txfm_kernel.cu).
// Block b: dequantised coefficients at coef + b*32*32 (row-major, min(w,32) columns per row),
// prediction/reconstruction at dst + b*w*h (row-major).
cudaError_t launch_inv_txfm_add(const int32_t* coef, uint16_t* dst, int n_blocks, int w, int h, int tx_type,
                                int bit_depth, cudaStream_t s);
// Batched encoder-side forward transform, every size and type (txfm_kernel.cu): block b reads resid + b*w*h (row-major) and
// writes coef + b * min(w,32) * min(h,32) (row-major, spec layout).
cudaError_t launch_fwd_txfm(const int16_t* resid, int32_t* coef, int n_blocks, int w, int h, int tx_type, cudaStream_t s);

// Source pyramid + hierarchical motion estimation (me_kernels.cu).  Search f of a launch matches the picture in
// slot cur_slot[f] against the SOURCE picture in slot ref_slot[f]: level l of slot k starts at cur[l] (ref[l]) +
// k * (elems0 >> 2l); strides are stride0 >> l.  mv2: scratch [n][n2y*n2x][2]; mv_out: [n][h8*w8][2].
constexpr int kMaxSearches = 64;
struct HmeLaunch {
  int32_t width, height, stride0;
  int32_t lambda;             // cost of one integer sample of deviation from the parent vector (SAD units, L0)
  int32_t shift2;             // bit depth - 8: the quarter-resolution search compares min(sample >> shift2, 255)
  size_t elems0;
  const uint16_t* cur[3];
  const uint16_t* ref[3];
  int16_t* mv2;
  int16_t* mv_out;
  uint8_t cur_slot[kMaxSearches], ref_slot[kMaxSearches];
  // vector-field regularisation (hme_sbrd kernels): lam_s per differing neighbour (0 = off), lam_r per bit of vector rate
  int32_t lam_s, lam_r, sbrd_passes;
  int16_t* mv_tmp;            // scratch [n][n1y*n1x][2]: the 16x16 block vectors
  uint32_t* hist;             // scratch [n][2][1024]: counts, largest key per bin
};
cudaError_t launch_pyramid(const uint16_t* l0, uint16_t* l1, uint16_t* l2, int stride0, int rows0, size_t elems0,
                           int n_frames, cudaStream_t s);
cudaError_t launch_hme(const HmeLaunch& p, int n_frames, cudaStream_t s);
// superblock-level rate-distortion sweeps over the vectors launch_hme left in mv_out (hist: [n * 2048 + n] words, mv_tmp: [n][n1][2])
cudaError_t launch_hme_sbrd(const HmeLaunch& p, int n, cudaStream_t s);

// Motion-compensated temporal filter of one key / anchor source picture (mctf_kernel.cu): weighted mean of the picture and
// up to kMaxNb neighbours in time, each compensated with mvs[k] (the picture searched against neighbour k).
constexpr int kMaxNb = 6;
struct MctfLaunch {
  Av1bGeom g;
  int32_t bit_depth, n_nb;
  int32_t thr_b, thr_p;       // block weight falls to zero at this luma mean squared error / sample weight at this squared difference
  const uint16_t* cur[3];
  uint16_t* out[3];
  const uint16_t* nb[kMaxNb][3];
  const int16_t* mvs[kMaxNb]; // [h8*w8][2] each
};
cudaError_t launch_mctf(const MctfLaunch& p, cudaStream_t s);
// Noise level of one source luma plane: hist[4096] (device) = histogram of the 16x16 blocks' sums of |I * N| >> 4
// (N = the 3x3 noise mask); the host takes the lower quartile outside bin 0 (capi_host.cc av1b_noise_from_hist).
cudaError_t launch_noise_hist(const Av1bGeom& g, const uint16_t* src_y, uint32_t* hist, cudaStream_t s);
// Scene-change scores of the n_frames luma planes at src_y (elems0 apart) against the picture before each (prev for the first;
// nullptr: no score): sum of absolute differences on the 1/8 x 1/8 sample grid from (4, 4).
cudaError_t launch_scene_score(const Av1bGeom& g, const uint16_t* src_y, size_t elems0, const uint16_t* prev, int n_frames,
                               uint32_t* score, cudaStream_t s);

// Inter frame encode (inter_kernel.cu): the n_frames frames of a launch share ONE reference picture and one
// quantiser (the frames between two anchors of the hierarchy); frame f of the launch has its source, outputs, block
// info and vectors at f * plane_elems[p] / f * map_elems from the given pointers.
struct InterLaunch {
  Av1bGeom g;
  int32_t bit_depth, base_q_idx, quant_rnd, dc_q, ac_q;
  int32_t n_frames;
  size_t plane_elems[3], map_elems;
  const uint16_t* src[3];
  const uint16_t* ref[3];     // reference frame after the in-loop filters (what the decoder holds)
  uint16_t* rec[3];
  int16_t* coef[3];
  Av1bBlockInfo* blocks;
  const uint8_t* part_map;    // 16x16 blocks, 8x8 at the picture edge (values 3 / 4)
  const int16_t* mvs;         // [h8*w8][2] (row, col), 1/8 luma samples
  int32_t tb_zero_thr;        // drop transform blocks whose levels sum to <= thr (16x16) / thr/2 (8x8)
  uint32_t dc_magic, ac_magic; // floor(2^32 / dc_q), floor(2^32 / ac_q): set by launch_inter_encode
  int32_t pack_levels;        // 1: transform blocks whose levels are all < 15 are stored as scan-ordered packed
                              //    symbols (sign | level | br ctx | base ctx) and flagged with bit 15 of eob
                              // 2: every coded transform block also gets its packed symbols in `digest` (token path)
  uint16_t* digest[3];        // pack_levels == 2: same layout as coef
  const uint8_t* qm[2] = {nullptr, nullptr};   // quantisation matrices as in IntraLaunch (a launch with one runs the <true> kernel)
};
cudaError_t launch_inter_encode(const InterLaunch& p, cudaStream_t s);
// Bottom-up merge of skipped inter siblings with equal vectors into 32x32 / 64x64 blocks (side info only).
cudaError_t launch_merge_skip(const Av1bGeom& g, Av1bBlockInfo* blocks, size_t map_elems, int n_frames, cudaStream_t s);

// Device tokenizer for inter frames (token_kernel.cu, tokens.h): one token per coded symbol, per tile in coding
// order.  All frames of a batch in one set of launches; key frames produce no tokens.
struct TokLaunch {
  Av1bGeom g;                    // inter-frame tile layout
  int32_t n_frames;
  uint64_t inter_mask;           // bit b: frame b of the batch is an inter frame
  int32_t cdef_bits;             // 0 when CDEF is off
  uint64_t nocdef_mask;          // bit b: frame b signals no CDEF (cdef_bits = 0 in its header: the non-reference frames)
  const Av1bBlockInfo* blocks;   // [n_frames][map_elems]
  const uint16_t* digest[3];     // [n_frames][plane_elems[p]]
  const int16_t* coef[3];
  const uint8_t* cdef_idx;       // [n_frames][nsb]
  size_t map_elems, plane_elems[3];
  uint8_t* mode_cls;             // scratch [n_frames][map_elems]
  uint32_t* blk_count;           // scratch [n_frames][map_elems]: tokens of the block whose origin the unit is
  uint32_t* sb_off;              // [n_frames * nsb + 1]: exclusive token offsets of the superblocks in coding order (+ total)
  const uint32_t* sb_of_order;   // [nsb]: coding order (tile by tile, raster inside a tile) -> superblock raster index
  const uint16_t* tile_of_sb;    // [nsb]
  uint32_t* tokens;
  uint32_t cap;                  // capacity of `tokens`
  const Av1bLrUnit* lr_units;    // [n_frames][lr_rows*lr_cols] luma restoration units, or nullptr
  int32_t lr_rows, lr_cols;
};
// Range coding of the token lists on the device (rc_kernel.cu): one warp per (frame, tile) of the batch's inter frames.
struct RcLaunch {
  int32_t n_frames, n_tiles, nsb;
  uint64_t inter_mask;
  const uint32_t* tokens;
  uint32_t tok_cap;              // capacity of `tokens`: tiles that end beyond it are left for the retry
  const uint32_t* sb_off;        // [n_frames * nsb + 1] token offsets (TokLaunch)
  const uint32_t* tile_first_k;  // [n_tiles + 1] first coding-order superblock of each tile
  const void* cdf_init;          // TileCdfs image of the frame's quantiser class (default CDFs)
  const void* cdf_init_alt;      // the image for the frames of alt_mask (non-reference frames: coarser quantiser)
  uint64_t alt_mask;
  uint8_t* region;               // scratch: tile (f, t) writes at 2 * (its first token) + 64 * (f * n_tiles + t)
  uint32_t* tile_len;            // [n_frames * n_tiles + 1]: byte counts, then (after the scan) offsets + total
  uint8_t* bytes;                // the batch's tile payloads, contiguous in (frame, tile) order
  uint32_t cap_bytes;
  uint32_t* overflow;            // set to 1 when a tile's region was too small
};
cudaError_t launch_rc(const RcLaunch& p, cudaStream_t s);
// in-place exclusive scan of v[0..n) on the device, v[n] = total
cudaError_t launch_scan_u32(uint32_t* v, int n, cudaStream_t s);
// mode classes + token counts + offsets (exclusive scan); then launch_tok_emit writes the tokens
cudaError_t launch_tok_count(const TokLaunch& p, cudaStream_t s);
cudaError_t launch_tok_emit(const TokLaunch& p, cudaStream_t s);

// Luma squared error and SSIM (8x8 windows) of n_frames reconstructions against their sources (quality_kernel.cu): reporting only.
struct QualityAcc { unsigned long long sse; float ssim_sum; uint32_t blocks; };
cudaError_t launch_quality(const Av1bGeom& g, int bit_depth, const uint16_t* rec_y, const uint16_t* src_y, size_t plane_elems,
                           QualityAcc* out, int n_frames, cudaStream_t s);

cudaError_t launch_deblock(const DeblockLaunch& p, int n_frames, cudaStream_t s);
cudaError_t launch_cdef(const CdefLaunch& p, int n_frames, cudaStream_t s);
cudaError_t launch_lr(const LrLaunch& p, int n_frames, cudaStream_t s);
// Per-unit choice among NONE / WIENER(cand) / SGRPROJ(cand) for the luma plane against the source: a candidate must
// beat NONE by more than `bias`.  units_out: [n_frames][unit_rows*unit_cols].
cudaError_t launch_lr_search(const LrLaunch& p, int n_frames, long long bias, Av1bLrUnit* units_out, cudaStream_t s);
cudaError_t launch_partition_fixed(const Av1bGeom& g, int blk_log2, uint8_t* map, int n_frames, cudaStream_t s);
// Key-frame partition by smoothness: 64x64 / 32x32 where the 4x4 box sums of the source stay within `thr` of a plane,
// else the fixed 16x16 blocks (8x8 at the picture edge).  map: [n_frames][map_elems].
cudaError_t launch_partition_smooth(const Av1bGeom& g, const uint16_t* src_y, size_t plane_elems, size_t map_elems, int thr,
                                    uint8_t* map, int n_frames, cudaStream_t s);
cudaError_t launch_intra_encode(const IntraLaunch& p, int n_frames, cudaStream_t s);
// Fast key-frame path for 16x16 blocks (8x8 at the picture edge): open-loop mode kernel + closed-loop
// reconstruction kernel + skip flags (intra_fast.cu).  Same results as launch_intra_encode.
cudaError_t launch_intra_fast(const IntraLaunch& p, int n_frames, cudaStream_t s);

}  // namespace av1b
