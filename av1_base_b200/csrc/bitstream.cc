// AV1 OBU packing + tile entropy coding on the host.  See bitstream.h.
// Section numbers refer to the AV1 Bitstream & Decoding Process Specification.
#include "bitstream.h"
#include <string.h>
#include <algorithm>
#include <thread>
#include <atomic>
#include "av1_tables.h"
#include "tokens.h"

namespace av1b {

// ------------------------------------------------------------------------------------------------
// Range encoder (the inverse of spec 8.2.6 "symbol decoding process")
// ------------------------------------------------------------------------------------------------
void RangeEncoder::finish(std::vector<uint8_t>& out) {
  // emit enough bits that any continuation decodes the same symbols, plus the terminating 1 bit
  uint32_t l = low_;
  int c = cnt_;
  int s = 10;
  const uint32_t m = 0x3FFF;
  uint32_t e = ((l + m) & ~m) | (m + 1);
  s += c;
  if (s > 0) {
    uint32_t n = (1u << (c + 16)) - 1;
    do {
      put((uint16_t)(e >> (c + 16)));
      e &= n;
      s -= 8;
      c -= 8;
      n >>= 8;
    } while (s > 0);
  }
  // carry propagation, last to first
  size_t base = out.size();
  out.resize(base + n_);
  uint32_t carry = 0;
  for (size_t i = n_; i-- > 0;) {
    carry += pre_[i];
    out[base + i] = (uint8_t)carry;
    carry >>= 8;
  }
}

// ------------------------------------------------------------------------------------------------
// OBU plumbing
// ------------------------------------------------------------------------------------------------
static void put_leb128(std::vector<uint8_t>& out, uint64_t v) {
  do {
    uint8_t b = v & 0x7F;
    v >>= 7;
    if (v) b |= 0x80;
    out.push_back(b);
  } while (v);
}

void append_obu(std::vector<uint8_t>& out, int obu_type, const std::vector<uint8_t>& payload) {
  out.push_back((uint8_t)((obu_type << 3) | 2));   // has_size_field = 1
  put_leb128(out, payload.size());
  out.insert(out.end(), payload.begin(), payload.end());
}

void write_temporal_delimiter(std::vector<uint8_t>& out) {
  std::vector<uint8_t> empty;
  append_obu(out, 2, empty);
}

static int bits_for(uint32_t v) { int n = 0; while (v) { n++; v >>= 1; } return n ? n : 1; }

void write_sequence_header(const Av1bSeqParams& seq, std::vector<uint8_t>& out) {
  BitWriter w;
  w.put(0, 3);   // seq_profile 0 (4:2:0, 8/10 bit)
  w.bit(0);      // still_picture
  w.bit(0);      // reduced_still_picture_header
  w.bit(0);      // timing_info_present_flag
  w.bit(0);      // initial_display_delay_present_flag
  w.put(0, 5);   // operating_points_cnt_minus_1
  w.put(0, 12);  // operating_point_idc[0]
  w.put(31, 5);  // seq_level_idx[0] = 31: "maximum parameters" (many-tile layouts exceed level limits)
  w.bit(0);      // seq_tier[0]
  int wb = bits_for(seq.width - 1), hb = bits_for(seq.height - 1);
  w.put(wb - 1, 4);
  w.put(hb - 1, 4);
  w.put(seq.width - 1, wb);
  w.put(seq.height - 1, hb);
  w.bit(0);   // frame_id_numbers_present_flag
  w.bit(0);   // use_128x128_superblock
  w.bit(0);   // enable_filter_intra
  w.bit(0);   // enable_intra_edge_filter
  w.bit(0);   // enable_interintra_compound
  w.bit(0);   // enable_masked_compound
  w.bit(0);   // enable_warped_motion
  w.bit(0);   // enable_dual_filter
  w.bit(0);   // enable_order_hint
  w.bit(0);   // seq_choose_screen_content_tools
  w.bit(0);   // seq_force_screen_content_tools = 0  (=> seq_force_integer_mv = SELECT, not coded)
  w.bit(0);   // enable_superres
  w.bit(seq.enable_cdef ? 1 : 0);
  w.bit(seq.enable_restoration ? 1 : 0);
  // color_config()
  w.bit(seq.bit_depth > 8);   // high_bitdepth
  w.bit(0);                   // mono_chrome
  if (seq.color_hdr) {
    w.bit(1);                 // color_description_present_flag
    w.put(9, 8);              // color_primaries  BT.2020
    w.put(16, 8);             // transfer_characteristics  SMPTE 2084 (PQ)
    w.put(9, 8);              // matrix_coefficients BT.2020 NCL
  } else {
    w.bit(0);
  }
  w.bit(0);                   // color_range (studio)
  w.put(0, 2);                // chroma_sample_position
  w.bit(0);                   // separate_uv_delta_q
  w.bit(seq.film_grain_present ? 1 : 0);   // film_grain_params_present
  w.trailing_bits();
  append_obu(out, 1, w.bytes());
}

// ------------------------------------------------------------------------------------------------
// Frame header (spec 5.9)
// ------------------------------------------------------------------------------------------------
// render_size() (spec 5.9.6): the size to present, when the source was padded to the coded size
static void write_render_size(const Av1bSeqParams& seq, BitWriter& w) {
  const bool diff = seq.render_width > 0 && seq.render_height > 0 && (seq.render_width != seq.width || seq.render_height != seq.height);
  w.bit(diff ? 1 : 0);         // render_and_frame_size_different
  if (diff) {
    w.put(seq.render_width - 1, 16);
    w.put(seq.render_height - 1, 16);
  }
}

static void write_frame_header(const Av1bSeqParams& seq, const Av1bFrameParams& fp, const Av1bGeom& g,
                               BitWriter& w) {
  const bool key = fp.frame_type == AV1B_KEY_FRAME;
  const bool inter = fp.frame_type == AV1B_INTER_FRAME;
  w.bit(0);                    // show_existing_frame
  w.put(fp.frame_type, 2);
  w.bit(1);                    // show_frame
  if (!key) w.bit(0);          // error_resilient_mode (implied 1 for shown key frames)
  w.bit(fp.disable_cdf_update ? 1 : 0);
  // allow_screen_content_tools = seq_force_screen_content_tools = 0 (not coded); force_integer_mv = 0
  w.bit(0);                    // frame_size_override_flag
  // order_hint: 0 bits (enable_order_hint = 0)
  if (inter) w.put(7, 3);      // primary_ref_frame = PRIMARY_REF_NONE: every frame starts from the default CDFs
  if (!key) w.put(inter ? (fp.non_reference ? 0x00 : 0xFF) : 0x01, 8);   // refresh_frame_flags (intra_only frames must not use 0xFF)
  if (inter) {
    // single reference design: all seven reference names point at slot 0 = the previous frame
    for (int i = 0; i < 7; i++) w.put(0, 3);   // ref_frame_idx[i]
    // frame_size(): from the sequence header; superres off
    write_render_size(seq, w);
    w.bit(0);                  // allow_high_precision_mv
    w.bit(0);                  // is_filter_switchable
    w.put(0, 2);               // interpolation_filter = EIGHTTAP (regular)
    w.bit(0);                  // is_motion_mode_switchable
    // use_ref_frame_mvs = 0 (enable_ref_frame_mvs = 0, not coded)
  } else {
    write_render_size(seq, w);
  }
  if (!fp.disable_cdf_update) w.bit(1);   // disable_frame_end_update_cdf
  // tile_info()
  {
    const int max_w_sb = 64, max_area_sb = 2304;
    int min_cols = av1b_tile_log2(max_w_sb, g.sb_cols);
    int max_cols = av1b_tile_log2(1, std::min(g.sb_cols, 64));
    int max_rows = av1b_tile_log2(1, std::min(g.sb_rows, 64));
    int min_tiles = std::max(min_cols, av1b_tile_log2(max_area_sb, g.sb_rows * g.sb_cols));
    w.bit(1);                  // uniform_tile_spacing_flag
    for (int k = min_cols; k < max_cols; k++) {
      if (k < g.tile_cols_log2) w.bit(1); else { w.bit(0); break; }
    }
    int min_rows = std::max(min_tiles - g.tile_cols_log2, 0);
    for (int k = min_rows; k < max_rows; k++) {
      if (k < g.tile_rows_log2) w.bit(1); else { w.bit(0); break; }
    }
    if (g.tile_cols_log2 > 0 || g.tile_rows_log2 > 0) {
      w.put(0, g.tile_cols_log2 + g.tile_rows_log2);   // context_update_tile_id
      w.put(3, 2);                                     // tile_size_bytes_minus_1
    }
  }
  // quantization_params()
  w.put(fp.base_q_idx, 8);
  w.bit(0);   // DeltaQYDc: delta_coded
  w.bit(0);   // DeltaQUDc
  w.bit(0);   // DeltaQUAc
  w.bit(fp.using_qmatrix ? 1 : 0);   // using_qmatrix
  if (fp.using_qmatrix) {
    w.put(fp.qm_level[0], 4);   // qm_y
    w.put(fp.qm_level[1], 4);   // qm_u; qm_v = qm_u (separate_uv_delta_q = 0)
  }
  w.bit(0);   // segmentation_enabled
  if (fp.base_q_idx > 0) w.bit(0);   // delta_q_present
  // loop_filter_params()   (base_q_idx > 0 => not CodedLossless)
  w.put(fp.lf_level[0], 6);
  w.put(fp.lf_level[1], 6);
  if (fp.lf_level[0] || fp.lf_level[1]) {
    w.put(fp.lf_level[2], 6);
    w.put(fp.lf_level[3], 6);
  }
  w.put(fp.lf_sharpness, 3);
  w.bit(0);   // loop_filter_delta_enabled
  // cdef_params()
  if (seq.enable_cdef) {
    w.put(fp.cdef_damping - 3, 2);
    w.put(fp.cdef_bits, 2);
    for (int i = 0; i < (1 << fp.cdef_bits); i++) {
      w.put(fp.cdef_y_strength[i] >> 2, 4);
      w.put(fp.cdef_y_strength[i] & 3, 2);
      w.put(fp.cdef_uv_strength[i] >> 2, 4);
      w.put(fp.cdef_uv_strength[i] & 3, 2);
    }
  }
  // lr_params()
  if (seq.enable_restoration) {
    static const int remap[4] = {0, 2, 3, 1};   // our RESTORE_* -> coded lr_type
    bool uses = false, uses_chroma = false;
    for (int p = 0; p < 3; p++) {
      w.put(remap[fp.lr_type[p]], 2);
      if (fp.lr_type[p] != AV1B_RESTORE_NONE) { uses = true; if (p) uses_chroma = true; }
    }
    if (uses) {
      w.bit(fp.lr_unit_shift > 0);
      if (fp.lr_unit_shift > 0) w.bit(fp.lr_unit_shift > 1);
      if (uses_chroma) w.bit(fp.lr_uv_shift);
    }
  }
  w.bit(0);   // tx_mode_select = 0 -> TX_MODE_LARGEST
  if (inter) w.bit(0);   // reference_select = 0 (single reference)
  // skip_mode_params: skip mode needs order hints (not coded); allow_warped_motion: enable_warped_motion = 0
  w.bit(0);   // reduced_tx_set
  if (inter) for (int i = 0; i < 7; i++) w.bit(0);   // global_motion_params(): is_global = 0 for LAST..ALTREF
  // film_grain_params() (spec 5.9.30; every frame is shown): white grain (no auto-regression) with one flat luma scaling
  // value, chroma scaled from luma -- what the temporal filter and the skipped blocks took out of the source comes back
  // as synthetic grain in the decoder (--film-grain, av1an.rs:14)
  if (seq.film_grain_present) {
    const int s = std::min(255, std::max(0, fp.grain_scaling));
    w.bit(s > 0);              // apply_grain
    if (s > 0) {
      w.put(fp.grain_seed & 0xFFFF, 16);
      if (inter) w.bit(1);     // update_grain: the parameters follow (no load from a reference)
      w.put(2, 4);             // num_y_points
      w.put(0, 8); w.put(s, 8); w.put(255, 8); w.put(s, 8);   // (point_y_value, point_y_scaling) x 2
      w.bit(1);                // chroma_scaling_from_luma  (num_cb_points = num_cr_points = 0)
      w.put(3, 2);             // grain_scaling_minus_8: noise = scaling * grain >> 11
      w.put(0, 2);             // ar_coeff_lag = 0: numPosLuma = 0, numPosChroma = 1
      w.put(128, 8);           // ar_coeffs_cb_plus_128[0]: no luma contribution
      w.put(128, 8);           // ar_coeffs_cr_plus_128[0]
      w.put(0, 2);             // ar_coeff_shift_minus_6
      w.put(0, 2);             // grain_scale_shift
      w.bit(1);                // overlap_flag
      w.bit(0);                // clip_to_restricted_range
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tile entropy coding (spec 5.11, coefficients 5.11.39)
// ------------------------------------------------------------------------------------------------
namespace {

void init_cdfs(TileCdfs& c, int base_q_idx) {
  const int q = base_q_idx <= 20 ? 0 : base_q_idx <= 60 ? 1 : base_q_idx <= 120 ? 2 : 3;
#define CP(dst, src) memcpy(dst, src, sizeof(dst))
  CP(c.partition, av1t_cdf_partition);
  CP(c.skip, av1t_cdf_skip);
  CP(c.kf_y_mode, av1t_cdf_kf_y_mode);
  CP(c.uv_mode, av1t_cdf_uv_mode);
  CP(c.angle_delta, av1t_cdf_angle_delta);
  CP(c.intra_ext_tx, av1t_cdf_intra_ext_tx);
  CP(c.txb_skip, av1t_cdf_txb_skip[q]);
  CP(c.eob_extra, av1t_cdf_eob_extra[q]);
  CP(c.dc_sign, av1t_cdf_dc_sign[q]);
  CP(c.eob_pt_16, av1t_cdf_eob_pt_16[q]);
  CP(c.eob_pt_32, av1t_cdf_eob_pt_32[q]);
  CP(c.eob_pt_64, av1t_cdf_eob_pt_64[q]);
  CP(c.eob_pt_128, av1t_cdf_eob_pt_128[q]);
  CP(c.eob_pt_256, av1t_cdf_eob_pt_256[q]);
  CP(c.eob_pt_512, av1t_cdf_eob_pt_512[q]);
  CP(c.eob_pt_1024, av1t_cdf_eob_pt_1024[q]);
  CP(c.coeff_base_eob, av1t_cdf_coeff_base_eob[q]);
  CP(c.coeff_base, av1t_cdf_coeff_base[q]);
  CP(c.coeff_br, av1t_cdf_coeff_br[q]);
  CP(c.cfl_sign, av1t_cdf_cfl_sign);
  CP(c.cfl_alpha, av1t_cdf_cfl_alpha);
  CP(c.switchable_restore, av1t_cdf_switchable_restore);
  CP(c.wiener_restore, av1t_cdf_wiener_restore);
  CP(c.sgrproj_restore, av1t_cdf_sgrproj_restore);
  CP(c.intra_inter, av1t_cdf_intra_inter);
  CP(c.single_ref, av1t_cdf_single_ref);
  CP(c.newmv, av1t_cdf_newmv); CP(c.zeromv, av1t_cdf_zeromv); CP(c.refmv, av1t_cdf_refmv); CP(c.drl, av1t_cdf_drl);
  CP(c.inter_ext_tx, av1t_cdf_inter_ext_tx);
  CP(c.y_mode, av1t_cdf_y_mode);
  CP(c.mv_joints, av1t_cdf_nmv_joints);
  CP(c.mvc[0].classes, av1t_cdf_nmv_c0_classes); CP(c.mvc[0].class0_fp, av1t_cdf_nmv_c0_class0_fp);
  CP(c.mvc[0].fp, av1t_cdf_nmv_c0_fp); CP(c.mvc[0].sign, av1t_cdf_nmv_c0_sign);
  CP(c.mvc[0].class0_hp, av1t_cdf_nmv_c0_class0_hp); CP(c.mvc[0].hp, av1t_cdf_nmv_c0_hp);
  CP(c.mvc[0].class0, av1t_cdf_nmv_c0_class0); CP(c.mvc[0].bits, av1t_cdf_nmv_c0_bits);
  CP(c.mvc[1].classes, av1t_cdf_nmv_c1_classes); CP(c.mvc[1].class0_fp, av1t_cdf_nmv_c1_class0_fp);
  CP(c.mvc[1].fp, av1t_cdf_nmv_c1_fp); CP(c.mvc[1].sign, av1t_cdf_nmv_c1_sign);
  CP(c.mvc[1].class0_hp, av1t_cdf_nmv_c1_class0_hp); CP(c.mvc[1].hp, av1t_cdf_nmv_c1_hp);
  CP(c.mvc[1].class0, av1t_cdf_nmv_c1_class0); CP(c.mvc[1].bits, av1t_cdf_nmv_c1_bits);
#undef CP
}

const uint8_t kIntraModeCtx[13] = {0, 1, 2, 3, 4, 4, 4, 4, 3, 0, 1, 2, 0};
// Mode_To_Txfm (spec): default transform type of an intra mode, used for chroma
const uint8_t kModeToTxfm[14] = {
    AV1B_DCT_DCT, AV1B_ADST_DCT, AV1B_DCT_ADST, AV1B_DCT_DCT, AV1B_ADST_ADST, AV1B_ADST_DCT,
    AV1B_DCT_ADST, AV1B_DCT_ADST, AV1B_ADST_DCT, AV1B_ADST_ADST, AV1B_ADST_DCT, AV1B_DCT_ADST,
    AV1B_ADST_ADST, AV1B_DCT_DCT};

enum { SET_DCTONLY = 0, SET_DCT_IDTX, SET_DTT4_IDTX, SET_DTT4_IDTX_1DDCT, SET_DTT9_IDTX_1DDCT, SET_ALL16 };

// intra transform set type for a square transform of log2 size n (2..6) (spec get_tx_set, intra)
inline int intra_tx_set_type(int tx_log2) {
  if (tx_log2 >= 5) return SET_DCTONLY;
  if (tx_log2 == 4) return SET_DTT4_IDTX;         // TX_SET_INTRA_2 (5 types)
  return SET_DTT4_IDTX_1DDCT;                     // TX_SET_INTRA_1 (7 types)
}
// inter transform set type (spec get_tx_set, is_inter = 1, reduced_tx_set = 0)
inline int inter_tx_set_type(int tx_log2) {
  if (tx_log2 >= 6) return SET_DCTONLY;
  if (tx_log2 == 5) return SET_DCT_IDTX;           // TX_SET_INTER_3
  if (tx_log2 == 4) return SET_DTT9_IDTX_1DDCT;    // TX_SET_INTER_2 (12 types)
  return SET_ALL16;                                 // TX_SET_INTER_1
}
inline int tx_class(int tx_type) {   // 0: 2D, 1: horizontal 1-D, 2: vertical 1-D
  switch (tx_type) {
    case AV1B_V_DCT: case AV1B_V_ADST: case AV1B_V_FLIPADST: return 2;
    case AV1B_H_DCT: case AV1B_H_ADST: case AV1B_H_FLIPADST: return 1;
    default: return 0;
  }
}

const int16_t* scan_for(int ns_log2, int tx_type) {
  const int cls = tx_class(tx_type);
  switch (ns_log2) {
    case 2: return cls == 2 ? av1t_scan_mrow_4x4 : cls == 1 ? av1t_scan_mcol_4x4 : av1t_scan_default_4x4;
    case 3: return cls == 2 ? av1t_scan_mrow_8x8 : cls == 1 ? av1t_scan_mcol_8x8 : av1t_scan_default_8x8;
    case 4: return cls == 2 ? av1t_scan_mrow_16x16 : cls == 1 ? av1t_scan_mcol_16x16 : av1t_scan_default_16x16;
    default: return av1t_scan_default_32x32;
  }
}
const int8_t* nz_offset_for(int ns_log2) {
  switch (ns_log2) {
    case 2: return av1t_nz_map_ctx_offset_4x4;
    case 3: return av1t_nz_map_ctx_offset_8x8;
    case 4: return av1t_nz_map_ctx_offset_16x16;
    default: return av1t_nz_map_ctx_offset_32x32;
  }
}

// Sub-exponential codes of the loop-restoration coefficients (spec 4.10.10 / 5.11.58), as literals of the range coder.
inline void lr_put_uniform(RangeEncoder& ec, int v, int n) {
  const int w = 32 - __builtin_clz((unsigned)n), m = (1 << w) - n;
  if (v < m) { ec.literal(v, w - 1); }
  else { ec.literal(m + ((v - m) >> 1), w - 1); ec.literal((v - m) & 1, 1); }
}
inline void lr_put_subexp(RangeEncoder& ec, int x, int num_syms, int k) {
  int i = 0, mk = 0;
  for (;;) {
    const int b2 = i ? k + i - 1 : k, a = 1 << b2;
    if (num_syms <= mk + 3 * a) { lr_put_uniform(ec, x - mk, num_syms - mk); return; }
    const int more = x >= mk + a;
    ec.literal(more, 1);
    if (more) { i++; mk += a; }
    else { ec.literal(x - mk, b2); return; }
  }
}
inline int lr_recenter(int r, int v) { return v > 2 * r ? v : (v >= r ? (v - r) << 1 : ((r - v) << 1) - 1); }
inline void lr_put_signed_subexp_with_ref(RangeEncoder& ec, int v, int low, int high, int k, int r) {
  const int mx = high - low, vv = v - low, rr = r - low;
  const int x = (rr << 1) <= mx ? lr_recenter(rr, vv) : lr_recenter(mx - 1 - rr, mx - 1 - vv);
  lr_put_subexp(ec, x, mx, k);
}
const int kLrTapMin[3] = {-5, -23, -17}, kLrTapMax[3] = {10, 8, 46}, kLrTapK[3] = {1, 2, 3};
const int kLrXqdMin[2] = {-96, -32}, kLrXqdMax[2] = {31, 95};

struct TileWriter {
  const Av1bSeqParams& seq;
  const Av1bFrameParams& fp;
  const Av1bGeom& g;
  const Av1bFrameSyms& sy;
  RangeEncoder ec;
  TileCdfs cdf;
  int mi_row_start, mi_row_end, mi_col_start, mi_col_end;
  // entropy contexts (4x4 units of each plane), above: indexed from tile start; left: within SB row
  std::vector<uint8_t> above_lvl[3], above_dc[3];
  uint8_t left_lvl[3][16], left_dc[3][16];
  uint8_t lvl_buf[(32 + 4) * (32 + 4) + 8];

  TileWriter(const Av1bSeqParams& s, const Av1bFrameParams& f, const Av1bGeom& gg, const Av1bFrameSyms& ss)
      : seq(s), fp(f), g(gg), sy(ss), ec(!f.disable_cdf_update) {}

  const Av1bBlockInfo& blk(int mi_row, int mi_col) const { return sy.blocks[(mi_row >> 1) * g.w8 + (mi_col >> 1)]; }
  bool avail_u(int r) const { return r > mi_row_start; }
  bool avail_l(int c) const { return c > mi_col_start; }

  void run(int tile_row, int tile_col, std::vector<uint8_t>& out) {
    mi_row_start = g.tile_row_start_sb[tile_row] * 16;
    mi_row_end = std::min(g.tile_row_start_sb[tile_row + 1] * 16, g.mi_rows);
    mi_col_start = g.tile_col_start_sb[tile_col] * 16;
    mi_col_end = std::min(g.tile_col_start_sb[tile_col + 1] * 16, g.mi_cols);
    init_cdfs(cdf, fp.base_q_idx);
    reset_lr_refs();
    if (fp.frame_type == AV1B_INTER_FRAME) {
      // compact tile-local copy of what motion vector prediction reads (8 bytes per 8x8 unit)
      t8c0 = mi_col_start >> 1; t8r0 = mi_row_start >> 1; t8w = (mi_col_end - mi_col_start + 1) >> 1;
      const int t8h = (mi_row_end - mi_row_start + 1) >> 1;
      units.resize((size_t)t8w * t8h);
      for (int y = 0; y < t8h; y++) {
        const Av1bBlockInfo* src = sy.blocks + (size_t)(t8r0 + y) * g.w8 + t8c0;
        Unit* dst = units.data() + (size_t)y * t8w;
        for (int x = 0; x < t8w; x++) {
          dst[x].mv[0] = src[x].mv[0]; dst[x].mv[1] = src[x].mv[1];
          if (dst[x].mv[0] & 1) dst[x].mv[0] += dst[x].mv[0] > 0 ? -1 : 1;   // lower_mv_precision (allow_high_precision_mv = 0)
          if (dst[x].mv[1] & 1) dst[x].mv[1] += dst[x].mv[1] > 0 ? -1 : 1;
          dst[x].w4 = (uint8_t)(1 << (src[x].blk_log2 - 2)); dst[x].is_inter = src[x].is_inter; dst[x].skip = src[x].skip; dst[x].coded = 0;
        }
      }
    }
    const int tw4 = mi_col_end - mi_col_start;
    for (int p = 0; p < 3; p++) {
      above_lvl[p].assign(tw4 + 32, 0);
      above_dc[p].assign(tw4 + 32, 0);
    }
    for (int r = mi_row_start; r < mi_row_end; r += 16) {
      memset(left_lvl, 0, sizeof(left_lvl));
      memset(left_dc, 0, sizeof(left_dc));
      for (int c = mi_col_start; c < mi_col_end; c += 16) {
        cdef_pending = true;   // clear_cdef(): cdef_idx is coded at the first non-skip block
        if (seq.enable_restoration) write_lr(r, c);
        partition(r, c, 6);
      }
    }
    ec.finish(out);
  }

  // ---- loop restoration unit syntax (spec 5.11.57 read_lr / 5.11.58 read_lr_unit) ----
  int ref_wiener[3][2][3], ref_sgr[3][2];
  void reset_lr_refs() {
    for (int p = 0; p < 3; p++) {
      for (int k = 0; k < 2; k++) { ref_wiener[p][k][0] = 3; ref_wiener[p][k][1] = -7; ref_wiener[p][k][2] = 15; }
      ref_sgr[p][0] = -32; ref_sgr[p][1] = 31;
    }
  }
  void write_lr(int r, int c) {
    for (int p = 0; p < 3; p++) {
      if (fp.lr_type[p] == AV1B_RESTORE_NONE) continue;
      const int ss = p > 0;
      int us = 64 << fp.lr_unit_shift;
      if (ss) us >>= fp.lr_uv_shift;
      const int unit_rows = std::max((((g.height + ss) >> ss) + (us >> 1)) / us, 1);
      const int unit_cols = std::max((((g.width + ss) >> ss) + (us >> 1)) / us, 1);
      const int m = 4 >> ss;
      const int row0 = (r * m + us - 1) / us, row1 = std::min(((r + 16) * m + us - 1) / us, unit_rows);
      const int col0 = (c * m + us - 1) / us, col1 = std::min(((c + 16) * m + us - 1) / us, unit_cols);
      for (int ur = row0; ur < row1; ur++)
        for (int uc = col0; uc < col1; uc++) {
          const Av1bLrUnit& u = sy.lr_units[p][ur * sy.lr_unit_cols[p] + uc];
          if (fp.lr_type[p] == AV1B_RESTORE_WIENER) ec.symbol(u.type == AV1B_RESTORE_WIENER, cdf.wiener_restore, 2);
          else if (fp.lr_type[p] == AV1B_RESTORE_SGRPROJ) ec.symbol(u.type == AV1B_RESTORE_SGRPROJ, cdf.sgrproj_restore, 2);
          else ec.symbol(u.type, cdf.switchable_restore, 3);   // NONE 0, WIENER 1, SGRPROJ 2
          if (u.type == AV1B_RESTORE_WIENER) {
            for (int pass = 0; pass < 2; pass++) {
              const int8_t* co = pass ? u.wiener_h : u.wiener_v;
              for (int j = p ? 1 : 0; j < 3; j++) {
                lr_put_signed_subexp_with_ref(ec, co[j], kLrTapMin[j], kLrTapMax[j] + 1, kLrTapK[j], ref_wiener[p][pass][j]);
                ref_wiener[p][pass][j] = co[j];
              }
            }
          } else if (u.type == AV1B_RESTORE_SGRPROJ) {
            ec.literal(u.sgr_set, 4);
            for (int i = 0; i < 2; i++) {
              const int radius = av1t_sgr_params[u.sgr_set][i];
              if (radius) {
                lr_put_signed_subexp_with_ref(ec, u.sgr_xqd[i], kLrXqdMin[i], kLrXqdMax[i] + 1, 4, ref_sgr[p][i]);
                ref_sgr[p][i] = u.sgr_xqd[i];
              } else {
                // not coded: decoder infers 0 (i == 0) or clip(128 - ref[0]) (i == 1)
                int v = 0;
                if (i == 1) v = std::min(std::max(128 - ref_sgr[p][0], kLrXqdMin[1]), kLrXqdMax[1]);
                ref_sgr[p][i] = v;
              }
            }
          }
        }
    }
  }

  // square-only partition tree; bl = log2 of block size in samples (6..3)
  void partition(int r, int c, int bl) {
    if (r >= g.mi_rows || c >= g.mi_cols) return;
    const int n4 = 1 << (bl - 2), half = n4 >> 1;
    const bool has_rows = (r + half) < g.mi_rows, has_cols = (c + half) < g.mi_cols;
    const Av1bBlockInfo& b = blk(r, c);
    const bool split = b.blk_log2 < bl;
    if (bl >= 3) {
      const int bsl = bl - 2;   // Mi_Width_Log2 of this size
      int above = avail_u(r) && (blk(r - 1, c).blk_log2 - 2) < bsl;
      int left = avail_l(c) && (blk(r, c - 1).blk_log2 - 2) < bsl;
      uint16_t* pc = cdf.partition[(bsl - 1) * 4 + left * 2 + above];
      const int nsym = bl == 3 ? 4 : 10;
      if (has_rows && has_cols) {
        ec.symbol(split ? 3 : 0, pc, nsym);
      } else if (has_cols || has_rows) {
        // split_or_horz / split_or_vert: binary symbol with probability gathered from the CDF
        // (spec 8.3.2, "psum").  We always choose split at frame edges.
        auto prob = [&](int k) -> int {   // P(partition == k) * 32768 from the inverted CDF
          int hi = k > 0 ? pc[k - 1] : 32768;
          return hi - pc[k];
        };
        int psum;
        if (has_cols) {   // !has_rows: split_or_horz, gather the "vertical alike" partitions
          psum = prob(2) + prob(3);
          if (bl != 3) psum += prob(4) + prob(6) + prob(7) + prob(9);   // HORZ_A VERT_A VERT_B VERT_4
        } else {          // split_or_vert, gather "horizontal alike"
          psum = prob(1) + prob(3);
          if (bl != 3) psum += prob(4) + prob(5) + prob(6) + prob(8);   // HORZ_A HORZ_B VERT_A HORZ_4
        }
        uint16_t tmp[16] = {(uint16_t)psum, 0, 0};   // icdf[0] = 32768 - P(not split) = psum (16 entries: the vector update)
        ec.symbol(1, tmp, 2);                        // derived CDF: its adaptation is discarded
      }
    }
    if (!split) {
      block(r, c, bl);
      return;
    }
    const int h = half;
    partition(r, c, bl - 1);
    partition(r, c + h, bl - 1);
    partition(r + h, c, bl - 1);
    partition(r + h, c + h, bl - 1);
  }

  // ---- inter frames (spec 5.11.18 inter_frame_mode_info, 7.10.2 motion vector prediction) ----
  // Per 8x8 unit of the tile: 0 = not coded yet, 1 = intra, 2 = inter without NEWMV, 3 = inter NEWMV
  struct Unit { int16_t mv[2]; uint8_t w4, is_inter, skip, coded; };
  std::vector<Unit> units;
  int t8c0 = 0, t8r0 = 0, t8w = 0;
  Unit& unit(int mi_r, int mi_c) { return units[(size_t)((mi_r >> 1) - t8r0) * t8w + ((mi_c >> 1) - t8c0)]; }
  uint8_t& coded(int mi_r, int mi_c) { return unit(mi_r, mi_c).coded; }
  bool is_inside(int mi_r, int mi_c) const {
    return mi_c >= mi_col_start && mi_c < mi_col_end && mi_r >= mi_row_start && mi_r < mi_row_end;
  }

  struct MvStack {
    int n = 0, num_new = 0, found = 0;
    int mv[8][2];
    int weight[8];
    int new_ctx = 0, ref_ctx = 0;
  };

  void add_ref_mv_candidate(MvStack& S, const Unit& cb, int weight) {
    if (!cb.is_inter) return;
    const int cand[2] = {cb.mv[0], cb.mv[1]};
    if (cb.coded == 3) S.num_new++;
    S.found = 1;
    int idx = 0;
    for (; idx < S.n; idx++) if (S.mv[idx][0] == cand[0] && S.mv[idx][1] == cand[1]) break;
    if (idx < S.n) S.weight[idx] += weight;
    else if (S.n < 8) { S.mv[S.n][0] = cand[0]; S.mv[S.n][1] = cand[1]; S.weight[S.n] = weight; S.n++; }
  }
  void scan_row(MvStack& S, int r, int c, int bw4, int delta_row) {
    int delta_col = 0;
    const int end4 = std::min(std::min(bw4, g.mi_cols - c), 16);
    if (std::abs(delta_row) > 1) { delta_row += r & 1; delta_col = 1 - (c & 1); }
    const bool step16 = bw4 >= 16;
    for (int i = 0; i < end4;) {
      const int mr = r + delta_row, mc = c + delta_col + i;
      if (!is_inside(mr, mc)) break;
      const Unit& cu = unit(mr, mc);
      int len = std::min(bw4, (int)cu.w4);
      if (std::abs(delta_row) > 1) len = std::max(2, len);
      if (step16) len = std::max(4, len);
      add_ref_mv_candidate(S, cu, len * 2);
      i += len;
    }
  }
  void scan_col(MvStack& S, int r, int c, int bh4, int delta_col) {
    int delta_row = 0;
    const int end4 = std::min(std::min(bh4, g.mi_rows - r), 16);
    if (std::abs(delta_col) > 1) { delta_row = 1 - (r & 1); delta_col += c & 1; }
    const bool step16 = bh4 >= 16;
    for (int i = 0; i < end4;) {
      const int mr = r + delta_row + i, mc = c + delta_col;
      if (!is_inside(mr, mc)) break;
      const Unit& cu = unit(mr, mc);
      int len = std::min(bh4, (int)cu.w4);
      if (std::abs(delta_col) > 1) len = std::max(2, len);
      if (step16) len = std::max(4, len);
      add_ref_mv_candidate(S, cu, len * 2);
      i += len;
    }
  }
  void scan_point(MvStack& S, int r, int c, int delta_row, int delta_col) {
    const int mr = r + delta_row, mc = c + delta_col;
    if (!is_inside(mr, mc)) return;
    const Unit& cu = unit(mr, mc);
    if (cu.coded != 0) add_ref_mv_candidate(S, cu, 4);
  }
  static void sort_stack(MvStack& S, int start, int end) {
    while (end > start) {
      int new_end = start;
      for (int idx = start + 1; idx < end; idx++) {
        if (S.weight[idx - 1] < S.weight[idx]) {
          std::swap(S.weight[idx - 1], S.weight[idx]);
          std::swap(S.mv[idx - 1][0], S.mv[idx][0]);
          std::swap(S.mv[idx - 1][1], S.mv[idx][1]);
          new_end = idx;
        }
      }
      end = new_end;
    }
  }
  // single reference (LAST_FRAME), no temporal candidates (use_ref_frame_mvs = 0), identity global motion
  void find_mv_stack(int r, int c, int bl, MvStack& S) {
    const int bw4 = 1 << (bl - 2), bh4 = bw4;
    S.found = 0;
    scan_row(S, r, c, bw4, -1);
    int found_above = S.found; S.found = 0;
    scan_col(S, r, c, bh4, -1);
    int found_left = S.found; S.found = 0;
    if (std::max(bw4, bh4) <= 16) scan_point(S, r, c, -1, bw4);
    if (S.found) found_above = 1;
    const int close_matches = found_above + found_left;
    const int num_nearest = S.n, num_new = S.num_new;
    for (int i = 0; i < num_nearest; i++) S.weight[i] += 640;   // REF_CAT_LEVEL
    S.found = 0;
    scan_point(S, r, c, -1, -1);
    if (S.found) found_above = 1;
    S.found = 0;
    scan_row(S, r, c, bw4, -3);
    if (S.found) found_above = 1;
    S.found = 0;
    scan_col(S, r, c, bh4, -3);
    if (S.found) found_left = 1;
    S.found = 0;
    scan_row(S, r, c, bw4, -5);
    if (S.found) found_above = 1;
    S.found = 0;
    scan_col(S, r, c, bh4, -5);
    if (S.found) found_left = 1;
    const int total_matches = found_above + found_left;
    sort_stack(S, 0, num_nearest);
    sort_stack(S, num_nearest, S.n);
    // extra search: every inter block uses the same reference, so row -1 / column -1 were already
    // collected above; the remaining entries up to two are the (zero) global motion vector
    for (int i = S.n; i < 2; i++) { S.mv[i][0] = 0; S.mv[i][1] = 0; S.weight[i] = 0; }
    if (close_matches == 0) { S.new_ctx = std::min(total_matches, 1); S.ref_ctx = total_matches; }
    else if (close_matches == 1) { S.new_ctx = 3 - std::min(num_new, 1); S.ref_ctx = 2 + total_matches; }
    else { S.new_ctx = 5 - std::min(num_new, 1); S.ref_ctx = 5; }
    // clamp (spec 7.10.2.14): MV_BORDER = 128 plus the block size, in 1/8 samples
    const int border_r = 128 + bh4 * 4 * 8, border_c = 128 + bw4 * 4 * 8;
    const int top = -(r * 4 * 8) - border_r, bottom = (g.mi_rows - bh4 - r) * 4 * 8 + border_r;
    const int left = -(c * 4 * 8) - border_c, right = (g.mi_cols - bw4 - c) * 4 * 8 + border_c;
    for (int i = 0; i < S.n; i++) {
      S.mv[i][0] = std::min(std::max(S.mv[i][0], top), bottom);
      S.mv[i][1] = std::min(std::max(S.mv[i][1], left), right);
    }
  }

  void write_mv_component(int comp, int diff) {
    TileCdfs::MvComp& m = cdf.mvc[comp];
    const int sign = diff < 0, mag = sign ? -diff : diff, offset = mag - 1;
    int cls = 0;
    if (offset >= 16) cls = 31 - __builtin_clz((unsigned)(offset >> 3));   // class c covers [2^(c+3), 2^(c+4))
    const int base = cls ? (2 << (cls + 2)) : 0;
    const int rem = offset - base, d = rem >> 3, fr = (rem >> 1) & 3;   // hp bit (rem & 1) is implied 1
    ec.symbol(sign, m.sign, 2);
    ec.symbol(cls, m.classes, 11);
    if (cls == 0) {
      ec.symbol(d, m.class0, 2);
      ec.symbol(fr, m.class0_fp[d], 4);
    } else {
      for (int i = 0; i < cls; i++) ec.symbol((d >> i) & 1, m.bits[i], 2);
      ec.symbol(fr, m.fp, 4);
    }
  }
  void write_mv(const int diff[2]) {
    const int joint = (diff[0] != 0 ? 2 : 0) | (diff[1] != 0 ? 1 : 0);   // ZERO, HNZVZ, HZVNZ, HNZVNZ
    ec.symbol(joint, cdf.mv_joints, 4);
    if (diff[0]) write_mv_component(0, diff[0]);
    if (diff[1]) write_mv_component(1, diff[1]);
  }

  void inter_block(int r, int c, int bl) {
    const Av1bBlockInfo& b = blk(r, c);
    const int n4 = 1 << (bl - 2);
    const bool au = avail_u(r), al = avail_l(c);
    {
      int ctx = (au ? blk(r - 1, c).skip : 0) + (al ? blk(r, c - 1).skip : 0);
      ec.symbol(b.skip ? 1 : 0, cdf.skip[ctx], 2);
    }
    if (!b.skip && seq.enable_cdef && fp.cdef_bits > 0) {
      if (cdef_pending) { ec.literal(sy.cdef_idx[(r >> 4) * g.sb_cols + (c >> 4)], fp.cdef_bits); cdef_pending = false; }
    }
    {
      const bool ai = au && !blk(r - 1, c).is_inter, li = al && !blk(r, c - 1).is_inter;
      int ctx;
      if (au && al) ctx = (li && ai) ? 3 : ((li || ai) ? 1 : 0);
      else if (au || al) ctx = 2 * (au ? (int)ai : (int)li);
      else ctx = 0;
      ec.symbol(b.is_inter ? 1 : 0, cdf.intra_inter[ctx], 2);
    }
    int mode_class = 1;
    if (b.is_inter) {
      // read_ref_frames(): LAST_FRAME = single_ref_p1 0, single_ref_p3 0, single_ref_p4 0
      const int cnt = (au && blk(r - 1, c).is_inter) + (al && blk(r, c - 1).is_inter);
      const int rctx = cnt == 0 ? 1 : 2;
      ec.symbol(0, cdf.single_ref[rctx][0], 2);
      ec.symbol(0, cdf.single_ref[rctx][2], 2);
      ec.symbol(0, cdf.single_ref[rctx][3], 2);
      MvStack S;
      find_mv_stack(r, c, bl, S);
      const int mv[2] = {b.mv[0], b.mv[1]};
      auto same = [&](int i) { return S.mv[i][0] == mv[0] && S.mv[i][1] == mv[1]; };
      int near_idx = -1;
      for (int i = 1; i < std::min(S.n, 4); i++) if (same(i)) { near_idx = i; break; }
      if (S.n > 0 && same(0)) {            // NEARESTMV
        ec.symbol(1, cdf.newmv[S.new_ctx], 2);
        ec.symbol(1, cdf.zeromv[0], 2);
        ec.symbol(0, cdf.refmv[S.ref_ctx], 2);
        mode_class = 2;
      } else if (near_idx > 0) {           // NEARMV, RefMvIdx = near_idx
        ec.symbol(1, cdf.newmv[S.new_ctx], 2);
        ec.symbol(1, cdf.zeromv[0], 2);
        ec.symbol(1, cdf.refmv[S.ref_ctx], 2);
        for (int idx = 1; idx < 3; idx++) {
          if (S.n > idx + 1) {
            const int dctx = S.weight[idx] >= 640 ? (S.weight[idx + 1] >= 640 ? 0 : 1) : 2;
            const int more = near_idx > idx;
            ec.symbol(more, cdf.drl[dctx], 2);
            if (!more) break;
          }
        }
        mode_class = 2;
      } else if (mv[0] == 0 && mv[1] == 0) {   // GLOBALMV (identity global motion)
        ec.symbol(1, cdf.newmv[S.new_ctx], 2);
        ec.symbol(0, cdf.zeromv[0], 2);
        mode_class = 2;
      } else {                             // NEWMV: predictor = the stack entry closest to mv
        ec.symbol(0, cdf.newmv[S.new_ctx], 2);
        int ref_idx = 0;
        long best = -1;
        const int n_pred = std::max(1, std::min(S.n, 3));
        for (int i = 0; i < n_pred; i++) {
          const long d = std::labs((long)mv[0] - S.mv[i][0]) + std::labs((long)mv[1] - S.mv[i][1]) + 4 * i;
          if (best < 0 || d < best) { best = d; ref_idx = i; }
        }
        for (int idx = 0; idx < 2; idx++) {
          if (S.n > idx + 1) {
            const int dctx = S.weight[idx] >= 640 ? (S.weight[idx + 1] >= 640 ? 0 : 1) : 2;
            const int more = ref_idx > idx;
            ec.symbol(more, cdf.drl[dctx], 2);
            if (!more) break;
          }
        }
        const int diff[2] = {mv[0] - S.mv[ref_idx][0], mv[1] - S.mv[ref_idx][1]};
        write_mv(diff);
        mode_class = 3;
      }
      // interpolation filter fixed, motion mode SIMPLE, no compound: nothing else is coded
    } else {
      // intra block inside an inter frame is not produced by this encoder
    }
    for (int y = 0; y < n4; y += 2) for (int x = 0; x < n4; x += 2) coded(r + y, c + x) = (uint8_t)mode_class;
    residual(r, c, bl, b, true);
  }

  void residual(int r, int c, int bl, const Av1bBlockInfo& b, bool is_inter) {
    const int n4 = 1 << (bl - 2);
    if (b.skip) {
      for (int p = 0; p < 3; p++) {
        const int ss = p > 0, x4 = (c - mi_col_start) >> ss, y4 = (r & 15) >> ss, n = std::max(1, n4 >> ss);
        memset(&above_lvl[p][x4], 0, n); memset(&above_dc[p][x4], 0, n);
        memset(&left_lvl[p][y4], 0, n); memset(&left_dc[p][y4], 0, n);
      }
      return;
    }
    for (int p = 0; p < 3; p++) {
      const int ss = p > 0;
      int tl = bl - ss;
      if (tl > 6) tl = 6;
      if (p > 0 && tl > 5) tl = 5;
      int tx_type = AV1B_DCT_DCT;
      if (p == 0) tx_type = b.tx_type_y;
      else if (is_inter) {
        // chroma of an inter block takes the luma transform type when the chroma set allows it
        tx_type = b.tx_type_y;
        if (!av1t_ext_tx_used[inter_tx_set_type(tl)][tx_type]) tx_type = AV1B_DCT_DCT;
      } else {
        tx_type = kModeToTxfm[b.uv_mode];
        if (!av1t_ext_tx_used[intra_tx_set_type(tl)][tx_type]) tx_type = AV1B_DCT_DCT;
      }
      coeffs(p, r, c, tl, b, tx_type, is_inter);
    }
  }

  void block(int r, int c, int bl) {
    if (fp.frame_type == AV1B_INTER_FRAME) { inter_block(r, c, bl); return; }
    const Av1bBlockInfo& b = blk(r, c);
    // intra_frame_mode_info()
    {
      int ctx = (avail_u(r) ? blk(r - 1, c).skip : 0) + (avail_l(c) ? blk(r, c - 1).skip : 0);
      ec.symbol(b.skip ? 1 : 0, cdf.skip[ctx], 2);
    }
    // read_cdef(): cdef_idx literal at the first non-skip block of each 64x64
    if (!b.skip && seq.enable_cdef && fp.cdef_bits > 0) {
      if (cdef_pending) { ec.literal(sy.cdef_idx[(r >> 4) * g.sb_cols + (c >> 4)], fp.cdef_bits); cdef_pending = false; }
    }
    {
      int am = avail_u(r) ? blk(r - 1, c).y_mode : AV1B_DC_PRED;
      int lm = avail_l(c) ? blk(r, c - 1).y_mode : AV1B_DC_PRED;
      ec.symbol(b.y_mode, cdf.kf_y_mode[kIntraModeCtx[am]][kIntraModeCtx[lm]], 13);
    }
    if (b.y_mode >= AV1B_V_PRED && b.y_mode <= AV1B_D67_PRED)
      ec.symbol(b.angle_y + 3, cdf.angle_delta[b.y_mode - AV1B_V_PRED], 7);
    {
      const int cfl_allowed = bl <= 5;
      ec.symbol(b.uv_mode, cdf.uv_mode[cfl_allowed][b.y_mode], cfl_allowed ? 14 : 13);
      if (b.uv_mode == AV1B_UV_CFL_PRED) {
        // cfl_alphas(): joint sign then magnitudes
        const int su = b.cfl_alpha_u >> 5 & 3, sv = b.cfl_alpha_v >> 5 & 3;   // 0 zero, 1 neg, 2 pos
        const int joint = su * 3 + sv - 1;
        ec.symbol(joint, cdf.cfl_sign, 8);
        if (su) ec.symbol(b.cfl_alpha_u & 15, cdf.cfl_alpha[(su - 1) * 3 + sv], 16);
        if (sv) ec.symbol(b.cfl_alpha_v & 15, cdf.cfl_alpha[(sv - 1) * 3 + su], 16);
      } else if (b.uv_mode >= AV1B_V_PRED && b.uv_mode <= AV1B_D67_PRED) {
        ec.symbol(b.angle_uv + 3, cdf.angle_delta[b.uv_mode - AV1B_V_PRED], 7);
      }
    }
    // TX_MODE_LARGEST: no tx_size syntax.
    residual(r, c, bl, b, false);
  }

  bool cdef_pending = false;

  // Coefficient syntax from the device-digested form: w[i] (scan order, i < eob) =
  // sign << 15 | min(|level|, 15) << 11 | br context << 6 | base context; all levels are < 15, so there
  // are no Golomb escapes.  Same symbols, in the same order, as the raster path below.
  void coeffs_packed(const uint16_t* w, int eob, int ns, int tx_ctx, int ptype, int plane, int x4, int y4, int abs_y4,
                     int w4, int max_x4, int max_y4, int* cul_out, int* dc_cat_out) {
    const int br_tx = std::min(tx_ctx, 3);
    for (int i = eob - 1; i >= 0; i--) {
      const unsigned v = w[i];
      const int level = (v >> 11) & 15;
      if (i == eob - 1) {
        const int c2 = i == 0 ? 0 : (i <= (ns * ns) / 8 ? 1 : (i <= (ns * ns) / 4 ? 2 : 3));
        ec.symbol(std::min(level, 3) - 1, cdf.coeff_base_eob[tx_ctx][ptype][c2], 3);
      } else {
        ec.symbol(std::min(level, 3), cdf.coeff_base[tx_ctx][ptype][v & 63], 4);
      }
      if (level > 2) {
        uint16_t* c = cdf.coeff_br[br_tx][ptype][(v >> 6) & 31];
        int rem = level - 3;
        for (int k = 0; k < 4; k++) {
          const int s = std::min(rem, 3);
          ec.symbol(s, c, 4);
          rem -= s;
          if (s < 3) break;
        }
      }
    }
    int cul = 0, dc_cat = 0;
    for (int i = 0; i < eob; i++) {
      const unsigned v = w[i];
      const int a = (v >> 11) & 15;
      if (!a) continue;
      const int sign = v >> 15;
      if (i == 0) {
        int dcs = 0;
        for (int k = 0; k < w4; k++) {
          if (x4 + k < max_x4) { int s = above_dc[plane][x4 + k]; dcs += s == 1 ? -1 : s == 2 ? 1 : 0; }
          if (abs_y4 + k < max_y4) { int s = left_dc[plane][y4 + k]; dcs += s == 1 ? -1 : s == 2 ? 1 : 0; }
        }
        ec.symbol(sign, cdf.dc_sign[ptype][dcs < 0 ? 1 : dcs > 0 ? 2 : 0], 2);
        dc_cat = sign ? 1 : 2;
      } else {
        ec.boolean(sign);
      }
      cul += a;
    }
    *cul_out = cul;
    *dc_cat_out = dc_cat;
  }

  // coefficient syntax for one transform block (one per plane per block in TX_MODE_LARGEST)
  void coeffs(int plane, int mi_r, int mi_c, int tl, const Av1bBlockInfo& b, int tx_type, bool is_inter) {
    const int ss = plane > 0;
    const int x0 = (mi_c * 4) >> ss, y0 = (mi_r * 4) >> ss;      // sample position in the plane
    const int n = 1 << tl, w4 = n >> 2;
    const int x4 = ((mi_c - mi_col_start) >> ss), y4 = (mi_r & 15) >> ss;
    const int max_x4 = ((g.mi_cols - mi_col_start) + ss) >> ss;   // relative to tile start
    const int max_y4 = (g.mi_rows + ss) >> ss;
    const int abs_y4 = mi_r >> ss;
    const int tx_ctx = tl - 2;                                    // square: txSzCtx = log2 - 2
    const int ptype = plane > 0;
    const int eob = b.eob[plane] & 0x7FFF;
    const bool packed = (b.eob[plane] & 0x8000) != 0;   // device-digested symbols (see coeffs_packed)
    // all_zero context (spec: get_txb_skip ctx)
    int ctx;
    if (plane == 0) {
      ctx = 0;   // transform covers the whole block
    } else {
      int above = 0, left = 0;
      for (int k = 0; k < w4; k++) {
        if (x4 + k < max_x4) above |= above_lvl[plane][x4 + k] | above_dc[plane][x4 + k];
        if (abs_y4 + k < max_y4) left |= left_lvl[plane][y4 + k] | left_dc[plane][y4 + k];
      }
      ctx = 7 + (above != 0) + (left != 0);
      // bw*bh > w*h only when the chroma block is larger than its transform (never here)
    }
    ec.symbol(eob == 0, cdf.txb_skip[tx_ctx][ctx], 2);
    int cul = 0, dc_cat = 0;
    if (eob > 0) {
      if (plane == 0 && is_inter) {
        const int set = inter_tx_set_type(tl);
        if (set != SET_DCTONLY && fp.base_q_idx > 0) {
          const int eset = set == SET_ALL16 ? 1 : (set == SET_DTT9_IDTX_1DDCT ? 2 : 3);
          const int nsym = set == SET_ALL16 ? 16 : (set == SET_DTT9_IDTX_1DDCT ? 12 : 2);
          ec.symbol(av1t_ext_tx_ind[set][tx_type], cdf.inter_ext_tx[eset][tl - 2], nsym);
        }
      } else if (plane == 0) {
        const int set = intra_tx_set_type(tl);
        if (set != SET_DCTONLY && fp.base_q_idx > 0) {
          const int eset = set == SET_DTT4_IDTX_1DDCT ? 1 : 2;
          const int nsym = set == SET_DTT4_IDTX_1DDCT ? 7 : 5;
          ec.symbol(av1t_ext_tx_ind[set][tx_type], cdf.intra_ext_tx[eset][tl - 2][b.y_mode], nsym);
        }
      }
      const int ns_log2 = std::min(tl, 5), ns = 1 << ns_log2;   // coded area is at most 32x32
      const int16_t* scan = scan_for(ns_log2, tx_type);
      const int cls = tx_class(tx_type);
      const int16_t* cf = sy.coef[plane] + av1b_coef_offset(g.sb_cols, plane, x0, y0);
      const int cs = ns;
      // eob position
      {
        int t = 0;
        static const int16_t start[13] = {0, 1, 2, 3, 5, 9, 17, 33, 65, 129, 257, 513, 1025};
        while (eob >= start[t + 1]) t++;     // eob in [start[t], start[t+1])
        // t is eobPt (1-based); symbol = t - 1
        const int eob_multi = 2 * ns_log2 - 4;
        const int ectx = cls == 0 ? 0 : 1;
        switch (eob_multi) {
          case 0: ec.symbol(t - 1, cdf.eob_pt_16[ptype][ectx], 5); break;
          case 1: ec.symbol(t - 1, cdf.eob_pt_32[ptype][ectx], 6); break;
          case 2: ec.symbol(t - 1, cdf.eob_pt_64[ptype][ectx], 7); break;
          case 3: ec.symbol(t - 1, cdf.eob_pt_128[ptype][ectx], 8); break;
          case 4: ec.symbol(t - 1, cdf.eob_pt_256[ptype][ectx], 9); break;
          case 5: ec.symbol(t - 1, cdf.eob_pt_512[ptype][ectx], 10); break;
          default: ec.symbol(t - 1, cdf.eob_pt_1024[ptype][ectx], 11); break;
        }
        const int nbits = t >= 3 ? t - 2 : 0;
        if (nbits > 0) {
          const int extra = eob - start[t];
          ec.symbol((extra >> (nbits - 1)) & 1, cdf.eob_extra[tx_ctx][ptype][t - 3], 2);
          for (int i = nbits - 2; i >= 0; i--) ec.boolean((extra >> i) & 1);
        }
      }
      if (packed) {
        coeffs_packed(reinterpret_cast<const uint16_t*>(cf), eob, ns, tx_ctx, ptype, plane, x4, y4, abs_y4, w4, max_x4, max_y4, &cul, &dc_cat);
      } else {
      // padded magnitude map for the neighbour contexts
      const int stride = ns + 4;
      memset(lvl_buf, 0, (size_t)stride * (ns + 4));
      for (int i = 0; i < eob; i++) {
        const int pos = scan[i], rr = pos >> ns_log2, cc = pos & (ns - 1);
        int v = cf[rr * cs + cc];
        v = v < 0 ? -v : v;
        lvl_buf[rr * stride + cc] = (uint8_t)std::min(v, 127);
      }
      const int8_t* nz_off = nz_offset_for(ns_log2);
      const int br_tx = std::min(tx_ctx, 3);
      for (int i = eob - 1; i >= 0; i--) {
        const int pos = scan[i], rr = pos >> ns_log2, cc = pos & (ns - 1);
        const uint8_t* L = lvl_buf + rr * stride + cc;
        const int level = L[0];
        if (i == eob - 1) {
          int c2 = i == 0 ? 0 : (i <= (ns * ns) / 8 ? 1 : (i <= (ns * ns) / 4 ? 2 : 3));
          ec.symbol(std::min(level, 3) - 1, cdf.coeff_base_eob[tx_ctx][ptype][c2], 3);
        } else {
          int mag;
          auto m3 = [](int v) { return v > 3 ? 3 : v; };
          if (cls == 0) mag = m3(L[1]) + m3(L[stride]) + m3(L[stride + 1]) + m3(L[2]) + m3(L[2 * stride]);
          else if (cls == 2) mag = m3(L[1]) + m3(L[stride]) + m3(L[2 * stride]) + m3(L[3 * stride]) + m3(L[4 * stride]);
          else mag = m3(L[1]) + m3(L[stride]) + m3(L[2]) + m3(L[3]) + m3(L[4]);
          int c2 = std::min((mag + 1) >> 1, 4);
          if (cls == 0) {
            c2 = (rr == 0 && cc == 0) ? 0 : c2 + nz_off[pos];
          } else {
            const int idx = cls == 2 ? rr : cc;
            c2 += 26 + 5 * std::min(idx, 2);
          }
          ec.symbol(std::min(level, 3), cdf.coeff_base[tx_ctx][ptype][c2], 4);
        }
        if (level > 2) {
          auto m15 = [](int v) { return v > 15 ? 15 : v; };
          int mag;
          if (cls == 0) mag = m15(L[1]) + m15(L[stride]) + m15(L[stride + 1]);
          else if (cls == 1) mag = m15(L[1]) + m15(L[2]) + m15(L[stride]);
          else mag = m15(L[stride]) + m15(L[2 * stride]) + m15(L[1]);
          mag = std::min((mag + 1) >> 1, 6);
          int c2;
          if (pos == 0) c2 = mag;
          else if (cls == 0) c2 = (rr < 2 && cc < 2) ? mag + 7 : mag + 14;
          else if (cls == 1) c2 = cc == 0 ? mag + 7 : mag + 14;
          else c2 = rr == 0 ? mag + 7 : mag + 14;
          int rem = level - 3;   // coded in up to 4 symbols of 0..3
          for (int k = 0; k < 4; k++) {
            const int s = std::min(rem, 3);
            ec.symbol(s, cdf.coeff_br[br_tx][ptype][c2], 4);
            rem -= s;
            if (s < 3) break;
          }
        }
      }
      // signs and Golomb remainders, forward scan order
      for (int i = 0; i < eob; i++) {
        const int pos = scan[i], rr = pos >> ns_log2, cc = pos & (ns - 1);
        int v = cf[rr * cs + cc];
        if (v == 0) continue;
        const int sign = v < 0;
        const int a = sign ? -v : v;
        if (i == 0) {
          int dcs = 0;
          for (int k = 0; k < w4; k++) {
            if (x4 + k < max_x4) { int s = above_dc[plane][x4 + k]; dcs += s == 1 ? -1 : s == 2 ? 1 : 0; }
            if (abs_y4 + k < max_y4) { int s = left_dc[plane][y4 + k]; dcs += s == 1 ? -1 : s == 2 ? 1 : 0; }
          }
          ec.symbol(sign, cdf.dc_sign[ptype][dcs < 0 ? 1 : dcs > 0 ? 2 : 0], 2);
          dc_cat = sign ? 1 : 2;
        } else {
          ec.boolean(sign);
        }
        if (a > 14) {
          const uint32_t x = (uint32_t)(a - 14);   // >= 1
          const int len = 32 - __builtin_clz(x);
          for (int k = 0; k < len - 1; k++) ec.boolean(0);
          for (int k = len - 1; k >= 0; k--) ec.boolean((x >> k) & 1);
        }
        cul += a;
      }
      }
      cul = std::min(cul, 63);
    }
    for (int k = 0; k < w4; k++) {
      above_lvl[plane][x4 + k] = (uint8_t)cul; above_dc[plane][x4 + k] = (uint8_t)dc_cat;
      left_lvl[plane][y4 + k] = (uint8_t)cul;  left_dc[plane][y4 + k] = (uint8_t)dc_cat;
    }
  }
};

}  // namespace

void pack_frame_header(const Av1bSeqParams& seq, const Av1bFrameParams& fp, const Av1bGeom& g, FramePack& fpk) {
  BitWriter hw;
  write_frame_header(seq, fp, g, hw);
  hw.byte_align();
  const int n_tiles = g.tile_cols * g.tile_rows;
  BitWriter tg;
  if (n_tiles > 1) tg.bit(0);   // tile_start_and_end_present_flag
  tg.byte_align();
  fpk.header = hw.bytes();
  fpk.header.insert(fpk.header.end(), tg.bytes().begin(), tg.bytes().end());
  fpk.tiles.assign(n_tiles, std::vector<uint8_t>());
}

void pack_tile(const Av1bSeqParams& seq, const Av1bFrameParams& fp, const Av1bGeom& g, const Av1bFrameSyms& syms,
               int tile, std::vector<uint8_t>& out) {
  TileWriter tw(seq, fp, g, syms);
  tw.run(tile / g.tile_cols, tile % g.tile_cols, out);
}

// Tile payload from a device-produced token list (tokens.h): the host keeps only the state that is serial by
// nature -- the range coder and the adapting CDFs.
namespace {
struct TokenCoder {
  TileCdfs cdf;
  RangeEncoder ec;
  uint16_t* base;
  int ref_wiener[3][2][3], ref_sgr[3][2];   // running references of the restoration coefficients (reset per tile)
  TokenCoder(const Av1bFrameParams& fp) : ec(!fp.disable_cdf_update), base(reinterpret_cast<uint16_t*>(&cdf)) {
    init_cdfs(cdf, fp.base_q_idx);
    for (int p = 0; p < 3; p++) {
      for (int k = 0; k < 2; k++) { ref_wiener[p][k][0] = 3; ref_wiener[p][k][1] = -7; ref_wiener[p][k][2] = 15; }
      ref_sgr[p][0] = -32; ref_sgr[p][1] = 31;
    }
  }
  __attribute__((always_inline)) inline void step(uint32_t t) {
    const uint32_t off = t & 0xFFFFu;
    if (off < TOK_FIRST_SPECIAL) {
      ec.symbol((int)(t >> 21), base + off, (int)((t >> 16) & 31));
    } else if (off == TOK_LR) {
      const int kind = (t >> 16) & 1, p = (t >> 17) & 3, pass = (t >> 19) & 1, j = (t >> 20) & 15, v = (int)(t >> 24) - 128;
      if (kind == 0) {
        lr_put_signed_subexp_with_ref(ec, v, kLrTapMin[j], kLrTapMax[j] + 1, kLrTapK[j], ref_wiener[p][pass][j]);
        ref_wiener[p][pass][j] = v;
      } else if (av1t_sgr_params[j][pass]) {      // j = parameter set, pass = weight index
        lr_put_signed_subexp_with_ref(ec, v, kLrXqdMin[pass], kLrXqdMax[pass] + 1, 4, ref_sgr[p][pass]);
        ref_sgr[p][pass] = v;
      } else {
        // radius 0: not coded, the decoder infers 0 (first weight) or clip(128 - ref[0]) (second)
        int w = 0;
        if (pass == 1) w = std::min(std::max(128 - ref_sgr[p][0], kLrXqdMin[1]), kLrXqdMax[1]);
        ref_sgr[p][pass] = w;
      }
    } else if (off == TOK_RAW) {
      ec.literal(t >> 21, (int)((t >> 16) & 31));
    } else if (off == TOK_GOLOMB) {
      const uint32_t x = t >> 16;
      const int len = 32 - __builtin_clz(x);
      for (int k = 0; k < len - 1; k++) ec.boolean(0);
      for (int k = len - 1; k >= 0; k--) ec.boolean((x >> k) & 1);
    } else {
      // forced split at the picture edge: binary symbol whose probability is gathered from the adaptive
      // partition CDF (spec 8.3.2); the derived CDF's adaptation is discarded
      const uint16_t* pc = cdf.partition[(t >> 16) & 31];
      const bool has_cols = (t >> 21) & 1, is8 = (t >> 22) & 1;
      auto prob = [&](int k) -> int { return (k > 0 ? pc[k - 1] : 32768) - pc[k]; };
      int psum;
      if (has_cols) { psum = prob(2) + prob(3); if (!is8) psum += prob(4) + prob(6) + prob(7) + prob(9); }
      else { psum = prob(1) + prob(3); if (!is8) psum += prob(4) + prob(5) + prob(6) + prob(8); }
      uint16_t tmp[16] = {(uint16_t)psum, 0, 0};   // 16 entries: the vector update of symbol()
      ec.symbol(1, tmp, 2);
    }
  }
};
}  // namespace

void tile_cdfs_default(int base_q_idx, void* dst) { init_cdfs(*static_cast<TileCdfs*>(dst), base_q_idx); }
size_t tile_cdfs_size() { return sizeof(TileCdfs); }

void pack_tile_tokens(const Av1bFrameParams& fp, const uint32_t* tok, size_t n, std::vector<uint8_t>& out) {
  TokenCoder c(fp);
  for (size_t i = 0; i < n; i++) c.step(tok[i]);
  c.ec.finish(out);
}

void assemble_frame(const FramePack& fpk, std::vector<uint8_t>& out) {
  size_t total = fpk.header.size();
  const int n_tiles = (int)fpk.tiles.size();
  for (auto& t : fpk.tiles) total += t.size() + 4;
  std::vector<uint8_t> payload;
  payload.reserve(total);
  payload = fpk.header;
  for (int t = 0; t < n_tiles; t++) {
    if (t != n_tiles - 1) {
      uint32_t sz = (uint32_t)fpk.tiles[t].size() - 1;   // tile_size_minus_1, 4 bytes LE
      for (int k = 0; k < 4; k++) payload.push_back((uint8_t)(sz >> (8 * k)));
    }
    payload.insert(payload.end(), fpk.tiles[t].begin(), fpk.tiles[t].end());
  }
  append_obu(out, 6, payload);
}

int write_frame(const Av1bSeqParams& seq, const Av1bFrameParams& fp, const Av1bGeom& g,
                const Av1bFrameSyms& syms, std::vector<uint8_t>& out, int n_threads) {
  FramePack fpk;
  pack_frame_header(seq, fp, g, fpk);
  const int n_tiles = (int)fpk.tiles.size();
  std::atomic<int> next(0);
  auto work = [&]() {
    for (;;) {
      int t = next.fetch_add(1);
      if (t >= n_tiles) break;
      pack_tile(seq, fp, g, syms, t, fpk.tiles[t]);
    }
  };
  if (n_threads <= 1 || n_tiles == 1) {
    work();
  } else {
    std::vector<std::thread> th;
    const int nt = std::min(n_threads, n_tiles);
    for (int i = 0; i < nt; i++) th.emplace_back(work);
    for (auto& t : th) t.join();
  }
  assemble_frame(fpk, out);
  return 0;
}

}  // namespace av1b
