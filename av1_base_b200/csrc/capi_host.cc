// C-ABI entry points that need no device: version, errors, bitstream packing.
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/av1b200.h"
#include "bitstream.h"
#include "capi_internal.h"
#include "av1_tables.h"
#include <math.h>
#include <algorithm>

namespace av1b {
thread_local std::string g_last_error;
void set_error(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}
}  // namespace av1b

extern "C" {

int av1b_version(char* buf, size_t cap) {
  static const char v[] = "av1b200 0.1.0 (AV1 encode backend, sm_100a)";
  if (!buf || cap == 0) return AV1B_ERR_INVALID;
  snprintf(buf, cap, "%s", v);
  return AV1B_OK;
}

// Frame-level filter parameters as a pure function of (bit depth, quantiser, frame type): deblocking
// levels and the eight CDEF strength presets the device chooses from per superblock.  The fits are
// the usual "pick from q" polynomials over the AC quantiser step.
int av1b_select_frame_params(int bit_depth, int base_q_idx, int frame_type, int loop_filters, Av1bFrameParams* fp) {
  if (!fp || base_q_idx < 0 || base_q_idx > 255 || (bit_depth != 8 && bit_depth != 10)) return AV1B_ERR_INVALID;
  memset(fp, 0, sizeof(*fp));
  fp->frame_type = frame_type;
  fp->base_q_idx = base_q_idx;
  fp->cdef_damping = 3;
  if (!loop_filters) return AV1B_OK;
  const int acq = bit_depth == 8 ? av1t_ac_q_8[base_q_idx] : av1t_ac_q_10[base_q_idx];
  const bool key = frame_type != AV1B_INTER_FRAME;
  long guess;
  if (bit_depth == 8) guess = key ? ((long)acq * 17563 - 421574 + (1 << 17)) >> 18 : ((long)acq * 6017 + 650707 + (1 << 17)) >> 18;
  else { guess = ((long)acq * 20723 + 4060632 + (1 << 19)) >> 20; if (key) guess -= 4; }
  const int lvl = (int)std::min<long>(63, std::max<long>(0, guess));
  fp->lf_level[0] = fp->lf_level[1] = fp->lf_level[2] = fp->lf_level[3] = lvl;
  fp->lf_sharpness = 0;
  const double q = (double)(acq >> (bit_depth - 8));
  auto cl = [](double v, int hi) { return (int)std::min<long>(hi, std::max<long>(0, lround(v))); };
  int y1, y2, u1, u2;
  if (key) {
    y1 = cl(q * q * 0.0000033731974 + q * 0.008070594 + 0.0187634, 15);
    y2 = cl(q * q * 0.0000029167343 + q * 0.0027798624 + 0.0079405, 3);
    u1 = cl(q * q * -0.0000130790995 + q * 0.012892405 - 0.00748388, 15);
    u2 = cl(q * q * 0.0000032651783 + q * 0.00035520183 + 0.00228092, 3);
  } else {
    y1 = cl(q * q * -0.0000023593946 + q * 0.0068615186 + 0.02709886, 15);
    y2 = cl(q * q * -0.00000057629734 + q * 0.0013993345 + 0.03831067, 3);
    u1 = cl(q * q * -0.0000007095069 + q * 0.0034628846 + 0.00887099, 15);
    u2 = cl(q * q * 0.00000023874085 + q * 0.00028223585 + 0.05576307, 3);
  }
  fp->cdef_damping = std::min(6, 3 + (base_q_idx >> 6));
  fp->cdef_bits = 3;
  const int P[8][4] = {
      {0, 0, 0, 0},
      {std::max(1, y1 / 2), 0, u1 / 2, 0},
      {y1, y2, u1, u2},
      {y1 / 2, y2, u1 / 2, u2},
      {std::min(15, y1 * 3 / 2 + 1), y2, std::min(15, u1 * 3 / 2), u2},
      {std::min(15, y1 * 2 + 1), std::min(3, y2 + 1), std::min(15, u1 * 2), u2},
      {y1, std::min(3, y2 + 1), u1, std::min(3, u2 + 1)},
      {std::min(15, y1 * 3 / 2 + 1), std::min(3, y2 + 1), std::min(15, u1 * 3 / 2), std::min(3, u2 + 1)},
  };
  for (int i = 0; i < 8; i++) {
    fp->cdef_y_strength[i] = P[i][0] * 4 + P[i][1];
    fp->cdef_uv_strength[i] = P[i][2] * 4 + P[i][3];
  }
  return AV1B_OK;
}

// Noise level out of the histogram noise_hist_kernel fills (oracle: orc_noise_estimate): centre of the bin that holds the
// lower quartile of the 16x16 blocks outside bin 0; 0 when all blocks are noiseless.
int av1b_noise_from_hist(const uint32_t* hist) {
  if (!hist) return 0;
  uint64_t n = 0;
  for (int b = 1; b < 4096; b++) n += hist[b];
  if (n == 0) return 0;
  const uint64_t want = (n + 3) / 4;
  uint64_t acc = 0;
  for (int b = 1; b < 4096; b++) {
    acc += hist[b];
    if (acc >= want) return (b << 4) + 8;
  }
  return (4095 << 4) + 8;
}

const char* av1b_last_error(void) { return av1b::g_last_error.c_str(); }

int av1b_pack_sequence_header(const Av1bSeqParams* seq, uint8_t* out, size_t cap, size_t* len) {
  if (!seq || !out || !len) { av1b::set_error("null argument"); return AV1B_ERR_INVALID; }
  std::vector<uint8_t> v;
  av1b::write_sequence_header(*seq, v);
  if (v.size() > cap) { av1b::set_error("output buffer too small"); return AV1B_ERR_NOMEM; }
  memcpy(out, v.data(), v.size());
  *len = v.size();
  return AV1B_OK;
}

int av1b_pack_frame(const Av1bSeqParams* seq, const Av1bFrameParams* fp, const Av1bFrameSyms* syms,
                    int n_threads, int with_td, uint8_t* out, size_t cap, size_t* len) {
  if (!seq || !fp || !syms || !out || !len) { av1b::set_error("null argument"); return AV1B_ERR_INVALID; }
  Av1bGeom g;
  if (av1b_geom_init(&g, seq->width, seq->height, fp->tile_cols_log2, fp->tile_rows_log2)) {
    av1b::set_error("unsupported frame size %dx%d (need multiples of 8, >= 16)", seq->width, seq->height);
    return AV1B_ERR_INVALID;
  }
  std::vector<uint8_t> v;
  if (with_td) av1b::write_temporal_delimiter(v);
  int rc = av1b::write_frame(*seq, *fp, g, *syms, v, n_threads);
  if (rc) { av1b::set_error("write_frame failed (%d)", rc); return AV1B_ERR_INTERNAL; }
  if (v.size() > cap) { av1b::set_error("output buffer too small"); return AV1B_ERR_NOMEM; }
  memcpy(out, v.data(), v.size());
  *len = v.size();
  return AV1B_OK;
}

// Same frame through the token path: tokens derived by the CPU statement of the device tokenizer, then
// range-coded tile by tile.  Inter frames only.  Must give the bytes of av1b_pack_frame.
int av1b_pack_frame_tokens(const Av1bSeqParams* seq, const Av1bFrameParams* fp, const Av1bFrameSyms* syms,
                           int with_td, uint8_t* out, size_t cap, size_t* len, uint64_t* n_tokens) {
  if (!seq || !fp || !syms || !out || !len) { av1b::set_error("null argument"); return AV1B_ERR_INVALID; }
  if (fp->frame_type != AV1B_INTER_FRAME) { av1b::set_error("token path codes inter frames only"); return AV1B_ERR_INVALID; }
  Av1bGeom g;
  if (av1b_geom_init(&g, seq->width, seq->height, fp->tile_cols_log2, fp->tile_rows_log2)) {
    av1b::set_error("unsupported frame size %dx%d", seq->width, seq->height);
    return AV1B_ERR_INVALID;
  }
  std::vector<std::vector<uint32_t>> toks;
  av1b::tokenize_frame_host(*fp, *seq, g, *syms, toks);
  av1b::FramePack fpk;
  av1b::pack_frame_header(*seq, *fp, g, fpk);
  uint64_t nt = 0;
  for (size_t t = 0; t < toks.size(); t++) { av1b::pack_tile_tokens(*fp, toks[t].data(), toks[t].size(), fpk.tiles[t]); nt += toks[t].size(); }
  std::vector<uint8_t> v;
  if (with_td) av1b::write_temporal_delimiter(v);
  av1b::assemble_frame(fpk, v);
  if (v.size() > cap) { av1b::set_error("output buffer too small"); return AV1B_ERR_NOMEM; }
  memcpy(out, v.data(), v.size());
  *len = v.size();
  if (n_tokens) *n_tokens = nt;
  return AV1B_OK;
}

}  // extern "C"
