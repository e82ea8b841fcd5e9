// Device-produced symbol streams for inter frames ("Tile/superblock entropy coding runs on the host over
// the device-produced symbol streams", BASELINE.json north_star).
//
// Everything in the AV1 tile syntax except the arithmetic coder state (range / low and the adaptive CDFs)
// is a pure function of the frame's block side information and quantised levels: which symbol is coded, with
// which CDF (the context), in which order.  The functions below derive that -- partition, skip, reference,
// motion vector prediction stack (spec 7.10.2) and mode, motion vector residual, transform-block contexts,
// end-of-block, coefficient and sign symbols -- and emit one 32-bit token per coded symbol:
//     bits  0..15  offset of the CDF inside TileCdfs (in uint16 units) | TOK_RAW | TOK_GOLOMB | TOK_PART_EDGE | TOK_LR
//     bits 16..20  number of symbols of that CDF              (TOK_RAW: number of literal bits, <= 11)
//     bits 21..31  symbol value                               (TOK_RAW: the literal, MSB first)
// The host then only walks the token list of a tile through the range coder (pack_tile_tokens, bitstream.cc).
// The same source is compiled for the device (token_kernel.cu: one warp per superblock, one lane per block)
// and for the host (tokenize_frame_host: the CPU statement of the same walk, used by the CPU tests to pin the
// token path against the block-walking tile writer of bitstream.cc bit for bit).
// Replaces work behind /root/reference/crates/daemon/src/encode/av1an.rs:126-139.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include "av1b_types.h"

namespace av1b {

// Adaptive CDF set of one tile (inverted CDFs, each vector followed by its terminator and counter).
struct TileCdfs {
  // ---- what inter frames code: the device range coder (rc_kernel.cu) keeps only this part in shared memory ----
  uint16_t partition[20][11];
  uint16_t skip[3][3];
  uint16_t txb_skip[5][13][3];
  uint16_t eob_extra[5][2][9][3];
  uint16_t dc_sign[2][3][3];
  uint16_t eob_pt_16[2][2][6];
  uint16_t eob_pt_32[2][2][7];
  uint16_t eob_pt_64[2][2][8];
  uint16_t eob_pt_128[2][2][9];
  uint16_t eob_pt_256[2][2][10];
  uint16_t eob_pt_512[2][2][11];
  uint16_t eob_pt_1024[2][2][12];
  uint16_t coeff_base_eob[5][2][4][4];
  uint16_t coeff_base[5][2][42][5];
  uint16_t coeff_br[5][2][21][5];
  uint16_t switchable_restore[4];
  uint16_t wiener_restore[3];
  uint16_t sgrproj_restore[3];
  uint16_t intra_inter[4][3];
  uint16_t single_ref[3][6][3];
  uint16_t newmv[6][3], zeromv[2][3], refmv[6][3], drl[3][3];
  uint16_t inter_ext_tx[4][4][17];
  uint16_t mv_joints[5];
  struct MvComp {
    uint16_t classes[12], class0_fp[2][5], fp[5], sign[3], class0_hp[3], hp[3], class0[3], bits[10][3];
  } mvc[2];
  uint16_t inter_end_[16];   // the SSE2 adaptation of the host coder may touch up to 16 entries from a CDF's start
  // ---- intra blocks only (key frames; intra blocks inside inter frames are not produced) ----
  uint16_t kf_y_mode[5][5][14];
  uint16_t uv_mode[2][13][15];
  uint16_t angle_delta[8][8];
  uint16_t intra_ext_tx[3][4][13][17];
  uint16_t cfl_sign[9];
  uint16_t cfl_alpha[6][17];
  uint16_t y_mode[4][14];
  uint16_t pad_[17];   // size multiple of 4 bytes; room behind the last CDF for the 16-entry vector update
};
static_assert(sizeof(TileCdfs) % 4 == 0, "TileCdfs is copied as 32-bit words");
static_assert(offsetof(TileCdfs, kf_y_mode) % 4 == 0, "the inter part is copied as 32-bit words");
static_assert(sizeof(TileCdfs) / 2 < 0xFFF0, "CDF offsets must fit 16 bits");

enum : uint32_t { TOK_RAW = 0xFFFFu, TOK_GOLOMB = 0xFFFEu, TOK_PART_EDGE = 0xFFFDu, TOK_LR = 0xFFFCu, TOK_FIRST_SPECIAL = TOK_LR };
// TOK_LR: one loop-restoration coefficient, coded by the host against its running reference (spec 5.11.58):
//   bit 16 kind (0 Wiener tap, 1 self-guided weight), bits 17..18 plane, bit 19 pass (Wiener: 0 vertical, 1 horizontal;
//   self-guided: weight index), bits 20..23 tap index j (Wiener) / parameter set (self-guided), bits 24..31 value + 128

#define AV1B_CDF_OFF(member) ((uint32_t)(offsetof(TileCdfs, member) / 2))

static inline AV1B_HD uint32_t tok_sym(uint32_t off, int nsym, int s) { return off | ((uint32_t)nsym << 16) | ((uint32_t)s << 21); }
static inline AV1B_HD uint32_t tok_raw(int nbits, uint32_t v) { return TOK_RAW | ((uint32_t)nbits << 16) | (v << 21); }
static inline AV1B_HD uint32_t tok_golomb(uint32_t x) { return TOK_GOLOMB | (x << 16); }
static inline AV1B_HD uint32_t tok_lr(int kind, int plane, int pass, int j_or_set, int value) {
  return TOK_LR | ((uint32_t)kind << 16) | ((uint32_t)plane << 17) | ((uint32_t)pass << 19) | ((uint32_t)j_or_set << 20) | ((uint32_t)(value + 128) << 24);
}

// What the token functions read of one inter frame and one tile.
struct TokFrame {
  const Av1bBlockInfo* blocks;   // [h8][w8], after merge_skip
  const uint8_t* mode_cls;       // [h8][w8]: 2 = inter block coded without NEWMV, 3 = NEWMV (tok_mode_class)
  const uint16_t* digest[3];     // per transform block, scan order: sign << 15 | min(|level|, 15) << 11 | br ctx << 6 | base ctx
  const int16_t* coef[3];        // raster levels (only read for |level| >= 15: Golomb remainder)
  const uint8_t* cdef_idx;       // [sb_rows][sb_cols]
  int32_t w8, h8, mi_cols, mi_rows, sb_cols;
  int32_t cdef_bits;             // 0 when CDEF is off
  const int16_t* scan[3];        // default scan of 4x4 / 8x8 / 16x16 (scan index -> raster position)
  const Av1bLrUnit* lr_units;    // luma restoration units [lr_rows][lr_cols] of 64x64 (frame type SWITCHABLE), or nullptr
  int32_t lr_rows, lr_cols;
  int32_t tx_sym_16, tx_sym_8;   // inter_ext_tx symbol of DCT_DCT in the 16x16 set (12 symbols) and the 8x8 / 4x4 set (16 symbols)
};
struct TokTile { int32_t mi_row_start, mi_row_end, mi_col_start, mi_col_end; };

struct TokSink {
  uint32_t* p;      // nullptr: count only
  uint32_t n, cap;
  AV1B_HD void put(uint32_t t) { if (p && n < cap) p[n] = t; n++; }
};

namespace tokdetail {

static inline AV1B_HD int imin(int a, int b) { return a < b ? a : b; }
static inline AV1B_HD int imax(int a, int b) { return a > b ? a : b; }
static inline AV1B_HD int iabs(int a) { return a < 0 ? -a : a; }

static inline AV1B_HD const Av1bBlockInfo& blk(const TokFrame& F, int mi_r, int mi_c) { return F.blocks[(size_t)(mi_r >> 1) * F.w8 + (mi_c >> 1)]; }
static inline AV1B_HD bool inside(const TokTile& T, int mi_r, int mi_c) {
  return mi_c >= T.mi_col_start && mi_c < T.mi_col_end && mi_r >= T.mi_row_start && mi_r < T.mi_row_end;
}
// Morton code of a 4x4 position inside its superblock (4 bits per axis)
static inline AV1B_HD unsigned morton4(int r, int c) {
  unsigned m = 0;
  for (int b = 0; b < 4; b++) m |= (((unsigned)c >> b) & 1u) << (2 * b) | (((unsigned)r >> b) & 1u) << (2 * b + 1);
  return m;
}
// Has the unit at (mr, mc) (inside the tile) been coded before the block whose origin is (r, c)?  Superblocks
// are walked in raster order inside the tile, blocks inside a superblock in Z order.
static inline AV1B_HD bool coded_before(int mr, int mc, int r, int c) {
  const int sr = mr >> 4, sc = mc >> 4, br = r >> 4, bc = c >> 4;
  if (sr != br) return sr < br;
  if (sc != bc) return sc < bc;
  return morton4(mr & 15, mc & 15) < morton4(r & 15, c & 15);
}

struct MvStack {
  int n, num_new, found;
  int mv[8][2];
  int weight[8];
  int new_ctx, ref_ctx;
};

static inline AV1B_HD int lower_mv(int v) { return (v & 1) ? v + (v > 0 ? -1 : 1) : v; }

static inline AV1B_HD void add_cand(const TokFrame& F, MvStack& S, int mr, int mc, int weight, bool count_new) {
  const Av1bBlockInfo& cb = blk(F, mr, mc);
  if (!cb.is_inter) return;
  const int c0 = lower_mv(cb.mv[0]), c1 = lower_mv(cb.mv[1]);
  if (count_new && F.mode_cls[(size_t)(mr >> 1) * F.w8 + (mc >> 1)] == 3) S.num_new++;
  S.found = 1;
  int idx = 0;
  for (; idx < S.n; idx++) if (S.mv[idx][0] == c0 && S.mv[idx][1] == c1) break;
  if (idx < S.n) S.weight[idx] += weight;
  else if (S.n < 8) { S.mv[S.n][0] = c0; S.mv[S.n][1] = c1; S.weight[S.n] = weight; S.n++; }
}
static inline AV1B_HD void scan_row(const TokFrame& F, const TokTile& T, MvStack& S, int r, int c, int bw4, int delta_row, bool cn) {
  int delta_col = 0;
  const int end4 = imin(imin(bw4, F.mi_cols - c), 16);
  const bool far = iabs(delta_row) > 1;
  if (far) { delta_row += r & 1; delta_col = 1 - (c & 1); }
  const bool step16 = bw4 >= 16;
  for (int i = 0; i < end4;) {
    const int mr = r + delta_row, mc = c + delta_col + i;
    if (!inside(T, mr, mc)) break;
    int len = imin(bw4, 1 << (blk(F, mr, mc).blk_log2 - 2));
    if (far) len = imax(2, len);
    if (step16) len = imax(4, len);
    add_cand(F, S, mr, mc, len * 2, cn);
    i += len;
  }
}
static inline AV1B_HD void scan_col(const TokFrame& F, const TokTile& T, MvStack& S, int r, int c, int bh4, int delta_col, bool cn) {
  int delta_row = 0;
  const int end4 = imin(imin(bh4, F.mi_rows - r), 16);
  const bool far = iabs(delta_col) > 1;
  if (far) { delta_row = 1 - (r & 1); delta_col += c & 1; }
  const bool step16 = bh4 >= 16;
  for (int i = 0; i < end4;) {
    const int mr = r + delta_row + i, mc = c + delta_col;
    if (!inside(T, mr, mc)) break;
    int len = imin(bh4, 1 << (blk(F, mr, mc).blk_log2 - 2));
    if (far) len = imax(2, len);
    if (step16) len = imax(4, len);
    add_cand(F, S, mr, mc, len * 2, cn);
    i += len;
  }
}
static inline AV1B_HD void scan_point(const TokFrame& F, const TokTile& T, MvStack& S, int r, int c, int delta_row, int delta_col, bool cn) {
  const int mr = r + delta_row, mc = c + delta_col;
  if (!inside(T, mr, mc)) return;
  if (coded_before(mr, mc, r, c)) add_cand(F, S, mr, mc, 4, cn);
}
static inline AV1B_HD void sort_stack(MvStack& S, int start, int end) {
  while (end > start) {
    int new_end = start;
    for (int idx = start + 1; idx < end; idx++) {
      if (S.weight[idx - 1] < S.weight[idx]) {
        int t = S.weight[idx - 1]; S.weight[idx - 1] = S.weight[idx]; S.weight[idx] = t;
        t = S.mv[idx - 1][0]; S.mv[idx - 1][0] = S.mv[idx][0]; S.mv[idx][0] = t;
        t = S.mv[idx - 1][1]; S.mv[idx - 1][1] = S.mv[idx][1]; S.mv[idx][1] = t;
        new_end = idx;
      }
    }
    end = new_end;
  }
}
// spec 7.10.2 for a single reference (LAST_FRAME), no temporal candidates, identity global motion.
// count_new = false leaves num_new / new_ctx undefined (mode pass: the stack entries do not depend on them).
static inline AV1B_HD void mv_stack(const TokFrame& F, const TokTile& T, int r, int c, int bl, bool cn, MvStack& S) {
  const int bw4 = 1 << (bl - 2), bh4 = bw4;
  S.n = 0; S.num_new = 0; S.found = 0; S.new_ctx = 0; S.ref_ctx = 0;
  scan_row(F, T, S, r, c, bw4, -1, cn);
  int found_above = S.found; S.found = 0;
  scan_col(F, T, S, r, c, bh4, -1, cn);
  int found_left = S.found; S.found = 0;
  if (imax(bw4, bh4) <= 16) scan_point(F, T, S, r, c, -1, bw4, cn);
  if (S.found) found_above = 1;
  const int close_matches = found_above + found_left;
  const int num_nearest = S.n, num_new = S.num_new;
  for (int i = 0; i < num_nearest; i++) S.weight[i] += 640;
  S.found = 0;
  scan_point(F, T, S, r, c, -1, -1, cn);
  if (S.found) found_above = 1;
  S.found = 0;
  scan_row(F, T, S, r, c, bw4, -3, cn);
  if (S.found) found_above = 1;
  S.found = 0;
  scan_col(F, T, S, r, c, bh4, -3, cn);
  if (S.found) found_left = 1;
  S.found = 0;
  scan_row(F, T, S, r, c, bw4, -5, cn);
  if (S.found) found_above = 1;
  S.found = 0;
  scan_col(F, T, S, r, c, bh4, -5, cn);
  if (S.found) found_left = 1;
  const int total_matches = found_above + found_left;
  sort_stack(S, 0, num_nearest);
  sort_stack(S, num_nearest, S.n);
  for (int i = S.n; i < 2; i++) { S.mv[i][0] = 0; S.mv[i][1] = 0; S.weight[i] = 0; }
  if (close_matches == 0) { S.new_ctx = imin(total_matches, 1); S.ref_ctx = total_matches; }
  else if (close_matches == 1) { S.new_ctx = 3 - imin(num_new, 1); S.ref_ctx = 2 + total_matches; }
  else { S.new_ctx = 5 - imin(num_new, 1); S.ref_ctx = 5; }
  const int border_r = 128 + bh4 * 4 * 8, border_c = 128 + bw4 * 4 * 8;
  const int top = -(r * 4 * 8) - border_r, bottom = (F.mi_rows - bh4 - r) * 4 * 8 + border_r;
  const int left = -(c * 4 * 8) - border_c, right = (F.mi_cols - bw4 - c) * 4 * 8 + border_c;
  for (int i = 0; i < S.n; i++) {
    S.mv[i][0] = imin(imax(S.mv[i][0], top), bottom);
    S.mv[i][1] = imin(imax(S.mv[i][1], left), right);
  }
}

// which of NEARESTMV (0) / NEARMV (near_idx 1..3) / GLOBALMV / NEWMV the block's vector maps to
struct ModeChoice { int kind; int idx; };   // kind: 0 nearest, 1 near (idx), 2 global, 3 new (idx = predictor)
static inline AV1B_HD ModeChoice choose_mode(const MvStack& S, int mv0, int mv1) {
  ModeChoice m; m.kind = 3; m.idx = 0;
  int near_idx = -1;
  for (int i = 1; i < imin(S.n, 4); i++) if (S.mv[i][0] == mv0 && S.mv[i][1] == mv1) { near_idx = i; break; }
  if (S.n > 0 && S.mv[0][0] == mv0 && S.mv[0][1] == mv1) { m.kind = 0; return m; }
  if (near_idx > 0) { m.kind = 1; m.idx = near_idx; return m; }
  if (mv0 == 0 && mv1 == 0) { m.kind = 2; return m; }
  long best = -1;
  const int n_pred = imax(1, imin(S.n, 3));
  for (int i = 0; i < n_pred; i++) {
    const long d = (long)iabs(mv0 - S.mv[i][0]) + (long)iabs(mv1 - S.mv[i][1]) + 4 * i;
    if (best < 0 || d < best) { best = d; m.idx = i; }
  }
  return m;
}

static inline AV1B_HD void put_mv_component(TokSink& K, int comp, int diff) {
  const uint32_t base = AV1B_CDF_OFF(mvc) + (uint32_t)comp * (uint32_t)(sizeof(TileCdfs::MvComp) / 2);
  const uint32_t o_classes = base + (uint32_t)(offsetof(TileCdfs::MvComp, classes) / 2);
  const uint32_t o_class0_fp = base + (uint32_t)(offsetof(TileCdfs::MvComp, class0_fp) / 2);
  const uint32_t o_fp = base + (uint32_t)(offsetof(TileCdfs::MvComp, fp) / 2);
  const uint32_t o_sign = base + (uint32_t)(offsetof(TileCdfs::MvComp, sign) / 2);
  const uint32_t o_class0 = base + (uint32_t)(offsetof(TileCdfs::MvComp, class0) / 2);
  const uint32_t o_bits = base + (uint32_t)(offsetof(TileCdfs::MvComp, bits) / 2);
  const int sign = diff < 0, mag = sign ? -diff : diff, offset = mag - 1;
  int cls = 0;
  if (offset >= 16) { int v = offset >> 3; while (v > 1) { v >>= 1; cls++; } }
  const int cbase = cls ? (2 << (cls + 2)) : 0;
  const int rem = offset - cbase, d = rem >> 3, fr = (rem >> 1) & 3;
  K.put(tok_sym(o_sign, 2, sign));
  K.put(tok_sym(o_classes, 11, cls));
  if (cls == 0) {
    K.put(tok_sym(o_class0, 2, d));
    K.put(tok_sym(o_class0_fp + (uint32_t)d * 5, 4, fr));
  } else {
    for (int i = 0; i < cls; i++) K.put(tok_sym(o_bits + (uint32_t)i * 3, 2, (d >> i) & 1));
    K.put(tok_sym(o_fp, 4, fr));
  }
}

// dc sign category (0 none, 1 negative, 2 positive) that the block covering (mi_r, mi_c) left in the entropy
// context of plane p
static inline AV1B_HD int dc_cat_of(const TokFrame& F, int p, int mi_r, int mi_c) {
  const Av1bBlockInfo& b = blk(F, mi_r, mi_c);
  if (b.skip || (b.eob[p] & 0x7FFF) == 0) return 0;
  const int ss = p > 0, n8 = 1 << (b.blk_log2 - 3);
  const int u_r = (mi_r >> 1) & ~(n8 - 1), u_c = (mi_c >> 1) & ~(n8 - 1);   // origin in 8x8 units
  const unsigned w = F.digest[p][av1b_coef_offset(F.sb_cols, p, (u_c * 8) >> ss, (u_r * 8) >> ss)];
  if (((w >> 11) & 15) == 0) return 0;
  return (w >> 15) ? 1 : 2;
}
static inline AV1B_HD bool nonzero_ctx_of(const TokFrame& F, int p, int mi_r, int mi_c) {
  const Av1bBlockInfo& b = blk(F, mi_r, mi_c);
  return !b.skip && (b.eob[p] & 0x7FFF) != 0;
}

}  // namespace tokdetail

// Pass 0: 2 / 3 = the inter block at (r, c) is coded without / with NEWMV (its neighbours' contexts need it).
static inline AV1B_HD int tok_mode_class(const TokFrame& F, const TokTile& T, int r, int c, int bl) {
  using namespace tokdetail;
  const Av1bBlockInfo& b = blk(F, r, c);
  MvStack S;
  mv_stack(F, T, r, c, bl, false, S);
  return choose_mode(S, b.mv[0], b.mv[1]).kind == 3 ? 3 : 2;
}

// Tokens that precede the blocks of the superblock at (sr, sc): the parameters of the luma restoration unit that
// starts in it (64x64 units: at most one per superblock).  Mirrors TileWriter::write_lr (bitstream.cc).
static inline AV1B_HD void tok_sb_lr(const TokFrame& F, int sr, int sc, TokSink& K) {
  if (!F.lr_units) return;
  const int ur = sr >> 4, uc = sc >> 4;
  if (ur >= F.lr_rows || uc >= F.lr_cols) return;
  const Av1bLrUnit& u = F.lr_units[ur * F.lr_cols + uc];
  K.put(tok_sym(AV1B_CDF_OFF(switchable_restore), 3, u.type));
  if (u.type == AV1B_RESTORE_WIENER) {
    for (int pass = 0; pass < 2; pass++)
      for (int j = 0; j < 3; j++) K.put(tok_lr(0, 0, pass, j, pass ? u.wiener_h[j] : u.wiener_v[j]));
  } else if (u.type == AV1B_RESTORE_SGRPROJ) {
    K.put(tok_raw(4, (uint32_t)u.sgr_set));
    for (int i = 0; i < 2; i++) K.put(tok_lr(1, 0, i, u.sgr_set, u.sgr_xqd[i]));
  }
}

// All tokens of the block whose origin is (r, c) (4x4 units), in coding order: partition symbols of the
// quadtree nodes that start at this block, block mode info, residual of the three planes.
// with_cdef: this is the first non-skip block of its superblock (it carries the cdef_idx literal).
static inline AV1B_HD void tok_block(const TokFrame& F, const TokTile& T, int r, int c, bool with_cdef, TokSink& K) {
  using namespace tokdetail;
  const Av1bBlockInfo& b = blk(F, r, c);
  const int bl = b.blk_log2;
  const bool au = r > T.mi_row_start, al = c > T.mi_col_start;
  // ---- partition symbols of every quadtree node whose first block this is ----
  for (int L = 6; L >= bl; L--) {
    const int n4 = 1 << (L - 2), half = n4 >> 1;
    if ((r & (n4 - 1)) || (c & (n4 - 1))) continue;
    const bool has_rows = (r + half) < F.mi_rows, has_cols = (c + half) < F.mi_cols;
    const bool split = L > bl;
    const int bsl = L - 2;
    const int above = au && (blk(F, r - 1, c).blk_log2 - 2) < bsl;
    const int left = al && (blk(F, r, c - 1).blk_log2 - 2) < bsl;
    const int pidx = (bsl - 1) * 4 + left * 2 + above;
    if (has_rows && has_cols) K.put(tok_sym(AV1B_CDF_OFF(partition) + (uint32_t)pidx * 11, L == 3 ? 4 : 10, split ? 3 : 0));
    else if (has_cols || has_rows) K.put(TOK_PART_EDGE | ((uint32_t)pidx << 16) | ((uint32_t)(has_cols ? 1 : 0) << 21) | ((uint32_t)(L == 3) << 22));
  }
  // ---- inter_frame_mode_info ----
  const Av1bBlockInfo* ba = au ? &blk(F, r - 1, c) : nullptr;
  const Av1bBlockInfo* bleft = al ? &blk(F, r, c - 1) : nullptr;
  K.put(tok_sym(AV1B_CDF_OFF(skip) + (uint32_t)((ba ? ba->skip : 0) + (bleft ? bleft->skip : 0)) * 3, 2, b.skip ? 1 : 0));
  if (!b.skip && with_cdef && F.cdef_bits > 0) K.put(tok_raw(F.cdef_bits, F.cdef_idx[(r >> 4) * F.sb_cols + (c >> 4)]));
  {
    const bool ai = ba && !ba->is_inter, li = bleft && !bleft->is_inter;
    int ctx;
    if (au && al) ctx = (li && ai) ? 3 : ((li || ai) ? 1 : 0);
    else if (au || al) ctx = 2 * (au ? (int)ai : (int)li);
    else ctx = 0;
    K.put(tok_sym(AV1B_CDF_OFF(intra_inter) + (uint32_t)ctx * 3, 2, 1));
  }
  {
    const int cnt = (ba && ba->is_inter) + (bleft && bleft->is_inter);
    const uint32_t o = AV1B_CDF_OFF(single_ref) + (uint32_t)(cnt == 0 ? 1 : 2) * 18;
    K.put(tok_sym(o + 0 * 3, 2, 0));
    K.put(tok_sym(o + 2 * 3, 2, 0));
    K.put(tok_sym(o + 3 * 3, 2, 0));
  }
  {
    MvStack S;
    mv_stack(F, T, r, c, bl, true, S);
    const int mv0 = b.mv[0], mv1 = b.mv[1];
    const ModeChoice m = choose_mode(S, mv0, mv1);
    const uint32_t o_new = AV1B_CDF_OFF(newmv) + (uint32_t)S.new_ctx * 3, o_zero = AV1B_CDF_OFF(zeromv);
    const uint32_t o_ref = AV1B_CDF_OFF(refmv) + (uint32_t)S.ref_ctx * 3, o_drl = AV1B_CDF_OFF(drl);
    if (m.kind == 0) {
      K.put(tok_sym(o_new, 2, 1)); K.put(tok_sym(o_zero, 2, 1)); K.put(tok_sym(o_ref, 2, 0));
    } else if (m.kind == 1) {
      K.put(tok_sym(o_new, 2, 1)); K.put(tok_sym(o_zero, 2, 1)); K.put(tok_sym(o_ref, 2, 1));
      for (int idx = 1; idx < 3; idx++) {
        if (S.n > idx + 1) {
          const int dctx = S.weight[idx] >= 640 ? (S.weight[idx + 1] >= 640 ? 0 : 1) : 2;
          const int more = m.idx > idx;
          K.put(tok_sym(o_drl + (uint32_t)dctx * 3, 2, more));
          if (!more) break;
        }
      }
    } else if (m.kind == 2) {
      K.put(tok_sym(o_new, 2, 1)); K.put(tok_sym(o_zero, 2, 0));
    } else {
      K.put(tok_sym(o_new, 2, 0));
      for (int idx = 0; idx < 2; idx++) {
        if (S.n > idx + 1) {
          const int dctx = S.weight[idx] >= 640 ? (S.weight[idx + 1] >= 640 ? 0 : 1) : 2;
          const int more = m.idx > idx;
          K.put(tok_sym(o_drl + (uint32_t)dctx * 3, 2, more));
          if (!more) break;
        }
      }
      const int d0 = mv0 - S.mv[m.idx][0], d1 = mv1 - S.mv[m.idx][1];
      K.put(tok_sym(AV1B_CDF_OFF(mv_joints), 4, (d0 != 0 ? 2 : 0) | (d1 != 0 ? 1 : 0)));
      if (d0) put_mv_component(K, 0, d0);
      if (d1) put_mv_component(K, 1, d1);
    }
  }
  if (b.skip) return;
  // ---- residual: one transform block per plane (TX_MODE_LARGEST, blocks of 16x16 or 8x8 here) ----
  for (int p = 0; p < 3; p++) {
    const int ss = p > 0, tl = bl - ss;            // log2 of the transform size (4 / 3 luma, 3 / 2 chroma)
    const int n = 1 << tl, w4 = n >> 2, tx_ctx = tl - 2, ptype = p > 0;
    const int eob = b.eob[p] & 0x7FFF;
    int ctx = 0;
    if (p > 0) {
      bool above = false, left = false;
      for (int k = 0; k < w4; k++) {
        if (au && c + (k << ss) < F.mi_cols) above |= nonzero_ctx_of(F, p, r - 1, c + (k << ss));
        if (al && r + (k << ss) < F.mi_rows) left |= nonzero_ctx_of(F, p, r + (k << ss), c - 1);
      }
      ctx = 7 + (above ? 1 : 0) + (left ? 1 : 0);
    }
    K.put(tok_sym(AV1B_CDF_OFF(txb_skip) + (uint32_t)(tx_ctx * 13 + ctx) * 3, 2, eob == 0));
    if (eob == 0) continue;
    if (p == 0) {
      if (tl == 4) K.put(tok_sym(AV1B_CDF_OFF(inter_ext_tx) + (uint32_t)(2 * 4 + (tl - 2)) * 17, 12, F.tx_sym_16));
      else K.put(tok_sym(AV1B_CDF_OFF(inter_ext_tx) + (uint32_t)(1 * 4 + (tl - 2)) * 17, 16, F.tx_sym_8));
    }
    {
      int t = 0;
      // eob in [start[t], start[t+1]) with start = 0,1,2,3,5,9,17,33,65,129,257
      while (eob >= (t + 1 <= 2 ? t + 1 : (1 << (t - 1)) + 1)) t++;
      const int start_t = t <= 2 ? t : (1 << (t - 2)) + 1;
      const int eob_multi = 2 * tl - 4;
      uint32_t o; int nsym;
      switch (eob_multi) {
        case 0: o = AV1B_CDF_OFF(eob_pt_16) + (uint32_t)(ptype * 2) * 6; nsym = 5; break;
        case 2: o = AV1B_CDF_OFF(eob_pt_64) + (uint32_t)(ptype * 2) * 8; nsym = 7; break;
        default: o = AV1B_CDF_OFF(eob_pt_256) + (uint32_t)(ptype * 2) * 10; nsym = 9; break;
      }
      K.put(tok_sym(o, nsym, t - 1));
      const int nbits = t >= 3 ? t - 2 : 0;
      if (nbits > 0) {
        const int extra = eob - start_t;
        K.put(tok_sym(AV1B_CDF_OFF(eob_extra) + (uint32_t)((tx_ctx * 2 + ptype) * 9 + (t - 3)) * 3, 2, (extra >> (nbits - 1)) & 1));
        if (nbits > 1) K.put(tok_raw(nbits - 1, (uint32_t)extra & ((1u << (nbits - 1)) - 1)));
      }
    }
    const int x0 = (c * 4) >> ss, y0 = (r * 4) >> ss;
    const size_t coff = av1b_coef_offset(F.sb_cols, p, x0, y0);
    const uint16_t* w = F.digest[p] + coff;
    const int br_tx = imin(tx_ctx, 3);
    bool escape = false;
    for (int i = eob - 1; i >= 0; i--) {
      const unsigned v = w[i];
      const int level = (v >> 11) & 15;
      if (i == eob - 1) {
        const int c2 = i == 0 ? 0 : (i <= (n * n) / 8 ? 1 : (i <= (n * n) / 4 ? 2 : 3));
        K.put(tok_sym(AV1B_CDF_OFF(coeff_base_eob) + (uint32_t)((tx_ctx * 2 + ptype) * 4 + c2) * 4, 3, imin(level, 3) - 1));
      } else {
        K.put(tok_sym(AV1B_CDF_OFF(coeff_base) + (uint32_t)((tx_ctx * 2 + ptype) * 42 + (int)(v & 63)) * 5, 4, imin(level, 3)));
      }
      if (level > 2) {
        const uint32_t o = AV1B_CDF_OFF(coeff_br) + (uint32_t)((br_tx * 2 + ptype) * 21 + (int)((v >> 6) & 31)) * 5;
        int rem = level - 3;
        for (int k = 0; k < 4; k++) {
          const int s = imin(rem, 3);
          K.put(tok_sym(o, 4, s));
          rem -= s;
          if (s < 3) break;
        }
        if (level == 15) escape = true;
      }
    }
    // signs (and Golomb remainders), forward scan order; runs of raw sign bits share one literal token
    uint32_t run = 0; int nrun = 0;
    for (int i = 0; i < eob; i++) {
      const unsigned v = w[i];
      const int a = (v >> 11) & 15;
      if (!a) continue;
      const int sign = v >> 15;
      if (i == 0) {
        int dcs = 0;
        for (int k = 0; k < w4; k++) {
          if (au && c + (k << ss) < F.mi_cols) { const int s = dc_cat_of(F, p, r - 1, c + (k << ss)); dcs += s == 1 ? -1 : s == 2 ? 1 : 0; }
          if (al && r + (k << ss) < F.mi_rows) { const int s = dc_cat_of(F, p, r + (k << ss), c - 1); dcs += s == 1 ? -1 : s == 2 ? 1 : 0; }
        }
        K.put(tok_sym(AV1B_CDF_OFF(dc_sign) + (uint32_t)(ptype * 3 + (dcs < 0 ? 1 : dcs > 0 ? 2 : 0)) * 3, 2, sign));
      } else {
        run = (run << 1) | (uint32_t)sign; nrun++;
      }
      bool flush = nrun == 11;
      if (escape && a == 15) {
        // true magnitude from the raster levels: position of scan index i
        const int pos = F.scan[tl - 2][i];
        int lv = F.coef[p][coff + pos];
        lv = lv < 0 ? -lv : lv;
        if (lv > 14) {
          if (nrun) { K.put(tok_raw(nrun, run)); run = 0; nrun = 0; }
          K.put(tok_golomb((uint32_t)(lv - 14)));
          flush = false;
        }
      }
      if (flush) { K.put(tok_raw(nrun, run)); run = 0; nrun = 0; }
    }
    if (nrun) K.put(tok_raw(nrun, run));
  }
}

}  // namespace av1b
