// C-ABI entry points that need no device: version, errors, bitstream packing.
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/av1b200.h"
#include "bitstream.h"
#include "capi_internal.h"

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

}  // extern "C"
