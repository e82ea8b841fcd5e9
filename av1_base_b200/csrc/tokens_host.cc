// CPU statement of the device tokenizer (token_kernel.cu): same token functions (tokens.h), walked
// sequentially.  Used by the CPU tests to pin the token path against the block-walking tile writer of
// bitstream.cc (identical bytes), and by av1b_pack_frame_tokens.  The product tokenizes on the device.
#include <string.h>
#include <algorithm>
#include <vector>
#include "av1_tables.h"
#include "bitstream.h"
#include "tokens.h"

namespace av1b {

namespace {

// scan-ordered packed words of one transform block from its raster levels, exactly what inter_encode_kernel
// stores: sign << 15 | min(|level|, 15) << 11 | br ctx << 6 | base ctx
void digest_tb(const int16_t* lev, int n, int eob, const int16_t* scan, const int8_t* nz_off, uint16_t* out) {
  const int s8 = n + 2;
  std::vector<uint8_t> lv8((size_t)(n + 2) * s8, 0);
  for (int r = 0; r < n; r++)
    for (int c = 0; c < n; c++) lv8[r * s8 + c] = (uint8_t)std::min(abs((int)lev[r * n + c]), 15);
  for (int si = 0; si < eob; si++) {
    const int pos = scan[si], t = pos / n, l = pos % n;
    const uint8_t* L = lv8.data() + t * s8 + l;
    const int m3 = std::min((int)L[1], 3) + std::min((int)L[s8], 3) + std::min((int)L[s8 + 1], 3) + std::min((int)L[2], 3) + std::min((int)L[2 * s8], 3);
    const int bctx = pos == 0 ? 0 : std::min((m3 + 1) >> 1, 4) + nz_off[pos];
    const int m15 = (int)L[1] + (int)L[s8] + (int)L[s8 + 1];
    const int brctx = std::min((m15 + 1) >> 1, 6) + (pos == 0 ? 0 : ((t < 2 && l < 2) ? 7 : 14));
    out[si] = (uint16_t)(((lev[pos] < 0 ? 1u : 0u) << 15) | ((unsigned)L[0] << 11) | ((unsigned)brctx << 6) | (unsigned)bctx);
  }
}

}  // namespace

void tokenize_frame_host(const Av1bFrameParams& fp, const Av1bSeqParams& seq, const Av1bGeom& g, const Av1bFrameSyms& sy,
                         std::vector<std::vector<uint32_t>>& tiles) {
  const size_t units = (size_t)g.w8 * g.h8;
  std::vector<uint16_t> digest[3];
  for (int p = 0; p < 3; p++) digest[p].assign((size_t)g.stride[p] * g.rows[p], 0);
  const int16_t* scans[3] = {av1t_scan_default_4x4, av1t_scan_default_8x8, av1t_scan_default_16x16};
  const int8_t* nzo[3] = {av1t_nz_map_ctx_offset_4x4, av1t_nz_map_ctx_offset_8x8, av1t_nz_map_ctx_offset_16x16};
  for (int uy = 0; uy < g.h8; uy++)
    for (int ux = 0; ux < g.w8; ux++) {
      const Av1bBlockInfo& b = sy.blocks[(size_t)uy * g.w8 + ux];
      const int n8 = 1 << (b.blk_log2 - 3);
      if ((ux | uy) & (n8 - 1)) continue;
      if (b.skip) continue;
      for (int p = 0; p < 3; p++) {
        const int ss = p > 0, tl = b.blk_log2 - ss, n = 1 << tl, eob = b.eob[p] & 0x7FFF;
        if (!eob) continue;
        const size_t off = av1b_coef_offset(g.sb_cols, p, (ux * 8) >> ss, (uy * 8) >> ss);
        digest_tb(sy.coef[p] + off, n, eob, scans[tl - 2], nzo[tl - 2], digest[p].data() + off);
      }
    }
  std::vector<uint8_t> mode_cls(units, 0);
  TokFrame F;
  F.blocks = sy.blocks; F.mode_cls = mode_cls.data();
  for (int p = 0; p < 3; p++) { F.digest[p] = digest[p].data(); F.coef[p] = sy.coef[p]; }
  F.cdef_idx = sy.cdef_idx;
  F.w8 = g.w8; F.h8 = g.h8; F.mi_cols = g.mi_cols; F.mi_rows = g.mi_rows; F.sb_cols = g.sb_cols;
  F.cdef_bits = seq.enable_cdef ? fp.cdef_bits : 0;
  for (int i = 0; i < 3; i++) F.scan[i] = scans[i];
  F.lr_units = (seq.enable_restoration && fp.lr_type[0] != AV1B_RESTORE_NONE) ? sy.lr_units[0] : nullptr;
  F.lr_rows = sy.lr_unit_rows[0]; F.lr_cols = sy.lr_unit_cols[0];
  F.tx_sym_16 = av1t_ext_tx_ind[4][AV1B_DCT_DCT]; F.tx_sym_8 = av1t_ext_tx_ind[5][AV1B_DCT_DCT];
  const int n_tiles = g.tile_cols * g.tile_rows;
  std::vector<TokTile> T(n_tiles);
  for (int t = 0; t < n_tiles; t++) {
    const int tr = t / g.tile_cols, tc = t % g.tile_cols;
    T[t].mi_row_start = g.tile_row_start_sb[tr] * 16; T[t].mi_row_end = std::min(g.tile_row_start_sb[tr + 1] * 16, g.mi_rows);
    T[t].mi_col_start = g.tile_col_start_sb[tc] * 16; T[t].mi_col_end = std::min(g.tile_col_start_sb[tc + 1] * 16, g.mi_cols);
  }
  auto tile_of = [&](int mi_r, int mi_c) {
    int tr = 0, tc = 0;
    while (g.tile_row_start_sb[tr + 1] * 16 <= mi_r) tr++;
    while (g.tile_col_start_sb[tc + 1] * 16 <= mi_c) tc++;
    return tr * g.tile_cols + tc;
  };
  // pass 0: mode class of every block
  for (int uy = 0; uy < g.h8; uy++)
    for (int ux = 0; ux < g.w8; ux++) {
      const Av1bBlockInfo& b = sy.blocks[(size_t)uy * g.w8 + ux];
      const int n8 = 1 << (b.blk_log2 - 3);
      if ((ux | uy) & (n8 - 1)) continue;
      const int cls = tok_mode_class(F, T[tile_of(uy * 2, ux * 2)], uy * 2, ux * 2, b.blk_log2);
      for (int yy = 0; yy < n8 && uy + yy < g.h8; yy++)
        for (int xx = 0; xx < n8 && ux + xx < g.w8; xx++) mode_cls[(size_t)(uy + yy) * g.w8 + ux + xx] = (uint8_t)cls;
    }
  // pass 1: tokens, tile by tile, superblocks in raster order, blocks in Z order
  tiles.assign(n_tiles, std::vector<uint32_t>());
  std::vector<uint32_t> buf(1 << 16);
  for (int t = 0; t < n_tiles; t++) {
    for (int sr = T[t].mi_row_start; sr < T[t].mi_row_end; sr += 16)
      for (int sc = T[t].mi_col_start; sc < T[t].mi_col_end; sc += 16) {
        bool cdef_pending = true;
        {
          TokSink K{buf.data(), 0, (uint32_t)buf.size()};
          tok_sb_lr(F, sr, sc, K);
          tiles[t].insert(tiles[t].end(), buf.begin(), buf.begin() + K.n);
        }
        for (int m = 0; m < 64; m++) {
          // Z order over the 8x8 units of the superblock
          const int ux = (m & 1) | ((m >> 1) & 2) | ((m >> 2) & 4), uy = ((m >> 1) & 1) | ((m >> 2) & 2) | ((m >> 3) & 4);
          const int r = sr + uy * 2, c = sc + ux * 2;
          if (r >= g.mi_rows || c >= g.mi_cols) continue;
          const Av1bBlockInfo& b = sy.blocks[(size_t)(r >> 1) * g.w8 + (c >> 1)];
          const int n8 = 1 << (b.blk_log2 - 3);
          if ((ux | uy) & (n8 - 1)) continue;
          const bool with_cdef = cdef_pending && !b.skip;
          if (with_cdef) cdef_pending = false;
          TokSink K{buf.data(), 0, (uint32_t)buf.size()};
          tok_block(F, T[t], r, c, with_cdef, K);
          tiles[t].insert(tiles[t].end(), buf.begin(), buf.begin() + K.n);
        }
      }
  }
}

}  // namespace av1b
