// Data interface between the device-side encode kernels and the host-side entropy coder /
// bitstream packer ("device-produced symbol streams", BASELINE.json north_star).
// Plain C structs: also read by the test oracle so that both sides can be compared buffer by buffer.
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { AV1B_KEY_FRAME = 0, AV1B_INTER_FRAME = 1, AV1B_INTRA_ONLY_FRAME = 2 };

// intra prediction modes (AV1 spec order)
enum {
  AV1B_DC_PRED = 0, AV1B_V_PRED, AV1B_H_PRED, AV1B_D45_PRED, AV1B_D135_PRED, AV1B_D113_PRED,
  AV1B_D157_PRED, AV1B_D203_PRED, AV1B_D67_PRED, AV1B_SMOOTH_PRED, AV1B_SMOOTH_V_PRED,
  AV1B_SMOOTH_H_PRED, AV1B_PAETH_PRED, AV1B_UV_CFL_PRED, AV1B_INTRA_MODES = 13
};
// transform types (AV1 spec order)
enum {
  AV1B_DCT_DCT = 0, AV1B_ADST_DCT, AV1B_DCT_ADST, AV1B_ADST_ADST, AV1B_FLIPADST_DCT,
  AV1B_DCT_FLIPADST, AV1B_FLIPADST_FLIPADST, AV1B_ADST_FLIPADST, AV1B_FLIPADST_ADST, AV1B_IDTX,
  AV1B_V_DCT, AV1B_H_DCT, AV1B_V_ADST, AV1B_H_ADST, AV1B_V_FLIPADST, AV1B_H_FLIPADST
};
enum { AV1B_RESTORE_NONE = 0, AV1B_RESTORE_WIENER = 1, AV1B_RESTORE_SGRPROJ = 2, AV1B_RESTORE_SWITCHABLE = 3 };

typedef struct Av1bSeqParams {
  int32_t width, height;        // luma samples of the coded frame
  int32_t bit_depth;            // 8 or 10
  int32_t enable_cdef;
  int32_t enable_restoration;
  int32_t fps_num, fps_den;
  int32_t color_hdr;            // 1: signal BT.2020 / PQ
  int32_t film_grain_present;   // 1: frame headers carry film grain parameters (--film-grain > 0)
  int32_t render_width, render_height;   // source size when it is not a multiple of 8 (the coded frame is the source padded by edge
                                //    replication; every frame header then carries render_size, spec 5.9.6); 0 = the coded size
} Av1bSeqParams;

typedef struct Av1bFrameParams {
  int32_t frame_type;           // AV1B_KEY_FRAME or AV1B_INTER_FRAME (single reference: the previous frame)
  int32_t base_q_idx;
  int32_t disable_cdf_update;   // 1: static default CDFs inside every tile
  int32_t tile_cols_log2, tile_rows_log2;
  int32_t lf_level[4];          // Y vertical edges, Y horizontal edges, U, V
  int32_t lf_sharpness;
  int32_t cdef_damping;         // 3..6
  int32_t cdef_bits;            // 0..3
  int32_t cdef_y_strength[8];   // pri * 4 + sec (sec in 0..3 as coded)
  int32_t cdef_uv_strength[8];
  int32_t lr_type[3];           // AV1B_RESTORE_* per plane
  int32_t lr_unit_shift;        // 0..2 : luma unit size 64 << shift
  int32_t lr_uv_shift;          // 0/1
  int32_t non_reference;        // 1: inter frame that updates no reference slot (refresh_frame_flags = 0): the frames after it
                                //    keep predicting from the last frame that did (one-level hierarchy, tools/rd_oracle.py --hier)
  int32_t grain_scaling;        // film grain synthesis (spec 5.9.30, only with seq.film_grain_present): 0 = apply_grain 0, else the
                                //    flat luma scaling value 1..255 (noise sigma = scaling / 64 in 8-bit units; chroma scaled from luma)
  int32_t grain_seed;           // 16 bits, varied from frame to frame
  int32_t using_qmatrix;        // 1: quantisation matrices (--enable-qm; spec 5.9.12 / 7.12.3) at the levels below
  int32_t qm_level[2];          // luma, chroma (qm_y; qm_u = qm_v): 0 (steepest) .. 15 (flat), from the quantiser index
} Av1bFrameParams;

// Frame geometry derived from (width, height): all in luma 4x4 "mode info" units unless noted.
typedef struct Av1bGeom {
  int32_t width, height;        // luma
  int32_t mi_cols, mi_rows;     // 4x4 units, even
  int32_t w8, h8;               // 8x8 units (the granularity of the per-block side info)
  int32_t sb_cols, sb_rows;     // 64x64 superblocks
  int32_t stride[3];            // sample stride of padded planes (multiple of 64 luma / 32 chroma)
  int32_t rows[3];              // allocated rows of padded planes
  int32_t tile_cols, tile_rows;
  int32_t tile_cols_log2, tile_rows_log2;
  int32_t tile_col_start_sb[65], tile_row_start_sb[65];
} Av1bGeom;

// Per-frame block side information, one entry per luma 8x8 unit (every unit of a block carries the
// block's values, so neighbour contexts can be read at any unit).
typedef struct Av1bBlockInfo {
  uint8_t blk_log2;   // 3..6 : block is (1<<blk_log2)^2 luma samples
  uint8_t y_mode;
  uint8_t uv_mode;
  uint8_t skip;
  int8_t angle_y;     // -3..3
  int8_t angle_uv;
  uint8_t tx_type_y;
  uint8_t cfl_alpha_u;  // reserved (CfL): signs/magnitudes packed
  uint16_t eob[3];    // per plane, valid at the block's top-left unit
  uint8_t cfl_alpha_v;
  uint8_t is_inter;   // 1: inter block predicted from LAST_FRAME with mv (inter frames only)
  int16_t mv[2];      // row, col in 1/8 luma samples (AV1 "Mv"), multiples of 2 (allow_high_precision_mv = 0)
} Av1bBlockInfo;        // 20 bytes

// Loop-restoration unit parameters
typedef struct Av1bLrUnit {
  int8_t type;          // AV1B_RESTORE_NONE / WIENER / SGRPROJ
  int8_t sgr_set;
  int8_t wiener_v[3], wiener_h[3];   // taps 0..2 (outer to inner), spec "LrWiener"
  int8_t sgr_xqd[2];
  int8_t pad[6];
} Av1bLrUnit;           // 16 bytes

// Host view of one coded frame's symbol streams.
typedef struct Av1bFrameSyms {
  const Av1bBlockInfo* blocks;   // [h8][w8]
  const int16_t* coef[3];        // quantised levels, transform-block contiguous (av1b_coef_offset):
                                 // level(row r, col c) of a block at offset + r * min(N,32) + c
  int32_t coef_stride[3];        // unused (kept for layout compatibility)
  const uint8_t* cdef_idx;       // [sb_rows][sb_cols]
  const Av1bLrUnit* lr_units[3]; // [unit_rows][unit_cols] per plane (may be NULL when lr_type==NONE)
  int32_t lr_unit_cols[3], lr_unit_rows[3];
} Av1bFrameSyms;

#ifdef __CUDACC__
#define AV1B_HD __host__ __device__
#else
#define AV1B_HD
#endif

// Quantised levels are stored transform-block contiguous: superblock-major, and inside a superblock
// in Morton order of its 8x8 luma units (4x4 chroma units), so that every aligned square block owns
// one contiguous range.  Returns the offset (in int16 elements) of the block whose top-left sample
// is (x, y) in `plane`; level (r, c) of that block is at offset + r * min(N, 32) + c.
// The per-plane element count equals that of the padded sample plane (stride * rows).
static inline AV1B_HD size_t av1b_coef_offset(int sb_cols, int plane, int x, int y) {
  const int ss = plane > 0, lsb = 6 - ss, lu = 3 - ss;
  const int ux = (x >> lu) & 7, uy = (y >> lu) & 7;
  const int m = (ux & 1) | ((uy & 1) << 1) | ((ux & 2) << 1) | ((uy & 2) << 2) | ((ux & 4) << 2) | ((uy & 4) << 3);
  const size_t sb = (size_t)(y >> lsb) * sb_cols + (x >> lsb);
  return (sb << (2 * lsb)) + ((size_t)m << (2 * lu));
}

static inline int av1b_tile_log2(int blk, int target) {
  int k = 0;
  while ((blk << k) < target) k++;
  return k;
}

// Fills the geometry; returns 0 or a negative error. Tiles: uniform spacing (AV1 spec 5.9.15).
static inline int av1b_geom_init(Av1bGeom* g, int width, int height, int tile_cols_log2, int tile_rows_log2) {
  if (width < 16 || height < 16 || (width & 7) || (height & 7) || width > 8192 || height > 4352) return -1;
  g->width = width; g->height = height;
  g->mi_cols = 2 * ((width + 7) >> 3); g->mi_rows = 2 * ((height + 7) >> 3);
  g->w8 = g->mi_cols >> 1; g->h8 = g->mi_rows >> 1;
  g->sb_cols = (g->mi_cols + 15) >> 4; g->sb_rows = (g->mi_rows + 15) >> 4;
  g->stride[0] = g->sb_cols * 64; g->stride[1] = g->stride[2] = g->sb_cols * 32;
  g->rows[0] = g->sb_rows * 64; g->rows[1] = g->rows[2] = g->sb_rows * 32;
  const int max_w_sb = 4096 >> 6, max_area_sb = (4096 * 2304) >> 12;
  int min_cols = av1b_tile_log2(max_w_sb, g->sb_cols);
  int max_cols = av1b_tile_log2(1, g->sb_cols < 64 ? g->sb_cols : 64);
  int max_rows = av1b_tile_log2(1, g->sb_rows < 64 ? g->sb_rows : 64);
  int min_tiles = av1b_tile_log2(max_area_sb, g->sb_rows * g->sb_cols);
  if (min_tiles < min_cols) min_tiles = min_cols;
  if (tile_cols_log2 < min_cols) tile_cols_log2 = min_cols;
  if (tile_cols_log2 > max_cols) tile_cols_log2 = max_cols;
  int min_rows = min_tiles - tile_cols_log2; if (min_rows < 0) min_rows = 0;
  if (tile_rows_log2 < min_rows) tile_rows_log2 = min_rows;
  if (tile_rows_log2 > max_rows) tile_rows_log2 = max_rows;
  g->tile_cols_log2 = tile_cols_log2; g->tile_rows_log2 = tile_rows_log2;
  int tw = (g->sb_cols + (1 << tile_cols_log2) - 1) >> tile_cols_log2, i = 0;
  for (int s = 0; s < g->sb_cols; s += tw) g->tile_col_start_sb[i++] = s;
  g->tile_col_start_sb[i] = g->sb_cols; g->tile_cols = i;
  int th = (g->sb_rows + (1 << tile_rows_log2) - 1) >> tile_rows_log2; i = 0;
  for (int s = 0; s < g->sb_rows; s += th) g->tile_row_start_sb[i++] = s;
  g->tile_row_start_sb[i] = g->sb_rows; g->tile_rows = i;
  return 0;
}

#ifdef __cplusplus
}
#endif
