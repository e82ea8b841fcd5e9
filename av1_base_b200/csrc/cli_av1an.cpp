// `av1an`-compatible command line front end of the B200 AV1 encode backend: the PATH executable the
// reference daemon execs (/root/reference/crates/daemon/src/encode/av1an.rs:79-107 build_av1an_command,
// :126-139 run_av1an; `av1an --version` in startup.rs:98-116).  Same argv, same exit-code contract
// (0 = success, non-zero = EncodeError::Av1anFailed(code)), complete file at -o or no file at all (a failed
// decode pipe or a failed audio copy is a failed job: the daemon replaces the source file with what it finds at -o),
// everything temporary under --temp: finished chunks wait there as packet files until the container is written.
//
//   av1an -i IN -o OUT --encoder svt-av1 --pix-format yuv420p10le --video-params "--crf 30 --preset 6 ..."
//         --audio-params "-c:a copy" --workers N --temp DIR
//
// --workers N is the number of GPUs this job uses: the clip is cut into closed-GOP chunks at detected
// scene cuts and at the latest every --keyint frames (av1an-style chunking), chunk c
// is encoded by worker c mod N on its own GPU, and the chunk streams are concatenated on the host in
// order (no exchange between GPUs).
// Input: YUV4MPEG2 (4:2:0, 8 or 10 bit) read directly and seekably; anything else is decoded through
// an `ffmpeg` pipe when ffmpeg is on PATH.  Output by extension: .mkv (Matroska, V_AV1), .ivf, .obu.
#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/file.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "../../include/av1b200.h"

namespace {

[[noreturn]] void die(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "av1an (av1b200): ");
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
  exit(code);
}

struct Options {
  std::string input, output, temp, encoder = "svt-av1", pix_format = "yuv420p10le", video_params, audio_params;
  int workers = 1;
  int crf = 30, preset = 6, keyint = 240, lookahead = -1, film_grain = 0, enable_qm = 0, qm_min = 8, qm_max = 15;   // SVT-AV1 defaults
  bool quiet = false;
  std::string mux_packets, mux_format;   // --mux-packets DIR --mux-format CWxCH,SWxSH,NUM:DEN,BITS (see main)
  bool no_scene_detection = false;   // --sc-method none / --no-scene-detection: split at --keyint only
  int min_scene_len = 12;            // --min-scene-len
};

// --video-params carries SVT-AV1 style flags (av1an.rs:14 SVT_PARAMS)
void parse_video_params(const std::string& vp, Options& o) {
  std::vector<std::string> tok;
  size_t i = 0;
  while (i < vp.size()) {
    while (i < vp.size() && isspace((unsigned char)vp[i])) i++;
    size_t j = i;
    while (j < vp.size() && !isspace((unsigned char)vp[j])) j++;
    if (j > i) tok.push_back(vp.substr(i, j - i));
    i = j;
  }
  for (size_t k = 0; k + 1 < tok.size(); k++) {
    const std::string& f = tok[k];
    const int v = atoi(tok[k + 1].c_str());
    if (f == "--crf") o.crf = v;
    else if (f == "--preset") o.preset = v;
    else if (f == "--keyint") o.keyint = v;
    else if (f == "--lookahead") o.lookahead = v;
    else if (f == "--film-grain") o.film_grain = v;
    else if (f == "--enable-qm") o.enable_qm = v;
    else if (f == "--qm-min") o.qm_min = v;
    else if (f == "--qm-max") o.qm_max = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Input: YUV4MPEG2
// ---------------------------------------------------------------------------------------------
struct Y4m {
  FILE* f = nullptr;
  bool pipe = false;
  int w = 0, h = 0, fps_num = 30, fps_den = 1, bits = 8;
  long header_len = 0;
  int64_t n_frames = -1;     // -1: unknown (pipe)
  size_t frame_bytes = 0;
};

bool y4m_parse_header(Y4m& y) {
  char line[512];
  int n = 0, c;
  while ((c = fgetc(y.f)) != EOF && c != '\n' && n < 511) line[n++] = (char)c;
  line[n] = 0;
  if (strncmp(line, "YUV4MPEG2", 9) != 0) return false;
  y.header_len = n + 1;
  std::string cs = "420";
  for (char* t = strtok(line + 9, " "); t; t = strtok(nullptr, " ")) {
    if (t[0] == 'W') y.w = atoi(t + 1);
    else if (t[0] == 'H') y.h = atoi(t + 1);
    else if (t[0] == 'F') sscanf(t + 1, "%d:%d", &y.fps_num, &y.fps_den);
    else if (t[0] == 'C') cs = t + 1;
  }
  if (cs.compare(0, 3, "420") != 0) die(3, "unsupported Y4M colourspace C%s (only 4:2:0)", cs.c_str());
  if (cs.find("p10") != std::string::npos) y.bits = 10;
  else if (cs.find("p12") != std::string::npos || cs.find("p16") != std::string::npos) die(3, "unsupported Y4M bit depth (C%s)", cs.c_str());
  const size_t bps = y.bits > 8 ? 2 : 1;
  y.frame_bytes = ((size_t)y.w * y.h + 2 * (size_t)((y.w + 1) / 2) * ((y.h + 1) / 2)) * bps;
  return y.w > 0 && y.h > 0;
}

// one frame as uint16 planes (contiguous Y, U, V); returns false at end of stream
bool y4m_read_frame(Y4m& y, std::vector<uint8_t>& raw, uint16_t* dst, int shift) {
  char hdr[256];
  int n = 0, c;
  while ((c = fgetc(y.f)) != EOF && c != '\n' && n < 255) hdr[n++] = (char)c;
  if (c == EOF) return false;
  hdr[n] = 0;
  if (strncmp(hdr, "FRAME", 5) != 0) die(3, "corrupt Y4M stream (expected FRAME)");
  if (y.bits > 8) {
    // little-endian 16-bit samples are already in our layout: read them in place
    if (fread(dst, 1, y.frame_bytes, y.f) != y.frame_bytes) die(3, "truncated Y4M frame");
  } else {
    raw.resize(y.frame_bytes);
    if (fread(raw.data(), 1, y.frame_bytes, y.f) != y.frame_bytes) die(3, "truncated Y4M frame");
    for (size_t i = 0; i < y.frame_bytes; i++) dst[i] = (uint16_t)(raw[i] << shift);
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// Output containers
// ---------------------------------------------------------------------------------------------
struct Packet { std::vector<uint8_t> data; bool key; };

void put_le(std::vector<uint8_t>& o, uint64_t v, int n) { for (int i = 0; i < n; i++) o.push_back((uint8_t)(v >> (8 * i))); }

// strips the temporal delimiter OBU (type 2) at the head of a temporal unit
const uint8_t* skip_td(const uint8_t* p, size_t& n) {
  if (n >= 2 && ((p[0] >> 3) & 15) == 2 && p[1] == 0) { n -= 2; return p + 2; }
  return p;
}
// first OBU of the given type inside a temporal unit (low-overhead format, obu_has_size_field = 1)
bool find_obu(const uint8_t* p, size_t n, int type, const uint8_t** start, size_t* len) {
  size_t i = 0;
  while (i < n) {
    const int t = (p[i] >> 3) & 15, ext = (p[i] >> 2) & 1;
    size_t j = i + 1 + ext, sz = 0;
    int sh = 0;
    while (j < n) { const uint8_t b = p[j++]; sz |= (size_t)(b & 0x7F) << sh; sh += 7; if (!(b & 0x80)) break; }
    if (t == type) { *start = p + i; *len = (j - i) + sz; return true; }
    i = j + sz;
  }
  return false;
}

void ebml_id(std::vector<uint8_t>& o, uint32_t id) {
  if (id > 0xFFFFFF) o.push_back((uint8_t)(id >> 24));
  if (id > 0xFFFF) o.push_back((uint8_t)(id >> 16));
  if (id > 0xFF) o.push_back((uint8_t)(id >> 8));
  o.push_back((uint8_t)id);
}
void ebml_size(std::vector<uint8_t>& o, uint64_t n) {   // always 8 bytes: simple and patchable
  o.push_back(0x01);
  for (int i = 6; i >= 0; i--) o.push_back((uint8_t)(n >> (8 * i)));
}
void ebml_uint(std::vector<uint8_t>& o, uint32_t id, uint64_t v) {
  ebml_id(o, id);
  int n = 1;
  while (n < 8 && (v >> (8 * n))) n++;
  o.push_back((uint8_t)(0x80 | n));
  for (int i = n - 1; i >= 0; i--) o.push_back((uint8_t)(v >> (8 * i)));
}
void ebml_bytes(std::vector<uint8_t>& o, uint32_t id, const void* p, size_t n) {
  ebml_id(o, id);
  ebml_size(o, n);
  o.insert(o.end(), (const uint8_t*)p, (const uint8_t*)p + n);
}
void ebml_str(std::vector<uint8_t>& o, uint32_t id, const char* s) { ebml_bytes(o, id, s, strlen(s)); }
void ebml_master(std::vector<uint8_t>& o, uint32_t id, const std::vector<uint8_t>& body) { ebml_bytes(o, id, body.data(), body.size()); }
void ebml_float(std::vector<uint8_t>& o, uint32_t id, double v) {
  ebml_id(o, id);
  o.push_back(0x88);
  uint64_t u;
  memcpy(&u, &v, 8);
  for (int i = 7; i >= 0; i--) o.push_back((uint8_t)(u >> (8 * i)));
}

// Packets of a finished chunk wait in <temp>/chunk_NNNNNN.pkt as [u32 size][u8 key][bytes]; the container is written
// by streaming them back in chunk order, so the memory a job needs does not grow with the length of the film.
struct PacketReader {
  std::vector<std::string> files;
  size_t next_file = 0;
  FILE* f = nullptr;
  bool next(Packet& p) {
    for (;;) {
      if (!f) {
        if (next_file >= files.size()) return false;
        f = fopen(files[next_file++].c_str(), "rb");
        if (!f) continue;     // a chunk without packets has no file
      }
      uint8_t h[5];
      if (fread(h, 1, 5, f) != 5) { fclose(f); f = nullptr; continue; }
      const uint32_t n = (uint32_t)h[0] | ((uint32_t)h[1] << 8) | ((uint32_t)h[2] << 16) | ((uint32_t)h[3] << 24);
      p.key = h[4] != 0;
      p.data.resize(n);
      if (fread(p.data.data(), 1, n, f) != n) { fclose(f); f = nullptr; return false; }
      return true;
    }
  }
  ~PacketReader() { if (f) fclose(f); }
};

bool put(FILE* f, const std::vector<uint8_t>& v) { return v.empty() || fwrite(v.data(), 1, v.size(), f) == v.size(); }

void ebml_uint_fixed8(std::vector<uint8_t>& o, uint32_t id, uint64_t v) {   // patchable: always 8 payload bytes
  ebml_id(o, id);
  o.push_back(0x88);
  for (int i = 7; i >= 0; i--) o.push_back((uint8_t)(v >> (8 * i)));
}
// SeekHead with the positions (relative to the Segment's data) of Info, Tracks and Cues; fixed size whatever the values
std::vector<uint8_t> mkv_seekhead(uint64_t info_pos, uint64_t tracks_pos, uint64_t cues_pos) {
  std::vector<uint8_t> sh;
  const uint32_t ids[3] = {0x1549A966, 0x1654AE6B, 0x1C53BB6B};
  const uint64_t pos[3] = {info_pos, tracks_pos, cues_pos};
  for (int k = 0; k < 3; k++) {
    std::vector<uint8_t> seek;
    const uint8_t idb[4] = {(uint8_t)(ids[k] >> 24), (uint8_t)(ids[k] >> 16), (uint8_t)(ids[k] >> 8), (uint8_t)ids[k]};
    ebml_bytes(seek, 0x53AB, idb, 4);
    ebml_uint_fixed8(seek, 0x53AC, pos[k]);
    ebml_master(sh, 0x4DBB, seek);
  }
  std::vector<uint8_t> out;
  ebml_master(out, 0x114D9B74, sh);
  return out;
}

// Minimal Matroska (video only): EBML header, Segment { SeekHead, Info, Tracks { V_AV1 }, Cluster* { SimpleBlock* }, Cues }.
// Clusters are written as they fill (one per key frame / 30 s); the Segment size, the Duration and the SeekHead are patched at
// the end, when the Cues (one point per cluster that starts with a key frame: what players seek by) have been written.
// w x h: the coded frame; show_w x show_h: the source size (smaller when the encoder padded it to multiples of 8)
bool write_mkv(const std::string& path, PacketReader& rd, int64_t* n_out, int w, int h, int show_w, int show_h, int fps_num, int fps_den, bool hbd) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  Packet pk;
  bool have = rd.next(pk);
  std::vector<uint8_t> out, hdr;
  ebml_uint(hdr, 0x4286, 1); ebml_uint(hdr, 0x42F7, 1); ebml_uint(hdr, 0x42F2, 4); ebml_uint(hdr, 0x42F3, 8);
  ebml_str(hdr, 0x4282, "matroska"); ebml_uint(hdr, 0x4287, 4); ebml_uint(hdr, 0x4285, 2);
  ebml_master(out, 0x1A45DFA3, hdr);
  ebml_id(out, 0x18538067);
  const long seg_size_pos = (long)out.size();
  ebml_size(out, 0);                                             // patched below
  const long seg_start = (long)out.size();
  const long seekhead_pos = (long)out.size();
  { const std::vector<uint8_t> ph = mkv_seekhead(0, 0, 0); out.insert(out.end(), ph.begin(), ph.end()); }   // patched below
  std::vector<uint8_t> info, tracks, te, video;
  const double frame_ms = 1000.0 * fps_den / fps_num;
  ebml_uint(info, 0x2AD7B1, 1000000);                           // TimestampScale: 1 ms
  const long dur_in_info = (long)info.size() + 2 + 1;           // id (2) + size byte, then 8 bytes of float
  ebml_float(info, 0x4489, 0.0);                                // Duration: patched below
  ebml_str(info, 0x4D80, "av1b200"); ebml_str(info, 0x5741, "av1b200");
  const long info_body_pos = (long)out.size() + 4 + 8;          // id (4) + 8-byte size
  const uint64_t info_seg_pos = (uint64_t)(out.size() - (size_t)seg_start);
  ebml_master(out, 0x1549A966, info);
  // CodecPrivate = AV1CodecConfigurationRecord: marker/version, profile/level, flags, then the sequence header OBU
  std::vector<uint8_t> av1c;
  const uint8_t* sh = nullptr; size_t shn = 0;
  if (have) find_obu(pk.data.data(), pk.data.size(), 1, &sh, &shn);
  av1c.push_back(0x81);
  av1c.push_back((uint8_t)((0 << 5) | 31));                       // seq_profile 0, seq_level_idx_0 31
  av1c.push_back((uint8_t)((0 << 7) | ((hbd ? 1 : 0) << 6) | (0 << 5) | (0 << 4) | (1 << 3) | (1 << 2) | 0));
  av1c.push_back(0);
  if (sh) av1c.insert(av1c.end(), sh, sh + shn);
  ebml_uint(video, 0xB0, (uint64_t)w); ebml_uint(video, 0xBA, (uint64_t)h);
  if (show_w != w || show_h != h) {
    // PixelCropBottom / PixelCropRight: the padding; DisplayWidth / DisplayHeight: what is left
    if (show_h != h) ebml_uint(video, 0x54AA, (uint64_t)(h - show_h));
    if (show_w != w) ebml_uint(video, 0x54DD, (uint64_t)(w - show_w));
    ebml_uint(video, 0x54B0, (uint64_t)show_w); ebml_uint(video, 0x54BA, (uint64_t)show_h);
  }
  ebml_uint(te, 0xD7, 1); ebml_uint(te, 0x73C5, 1); ebml_uint(te, 0x83, 1); ebml_uint(te, 0x9C, 0);
  ebml_str(te, 0x86, "V_AV1");
  ebml_bytes(te, 0x63A2, av1c.data(), av1c.size());
  ebml_uint(te, 0x23E383, (uint64_t)(frame_ms * 1e6));          // DefaultDuration (ns)
  ebml_master(te, 0xE0, video);
  ebml_master(tracks, 0xAE, te);
  const uint64_t tracks_seg_pos = (uint64_t)(out.size() - (size_t)seg_start);
  ebml_master(out, 0x1654AE6B, tracks);
  bool ok = put(f, out);
  uint64_t seg_bytes = out.size() - (size_t)seg_start;
  // clusters: a new one at every key frame and at least every 30000 ms of timestamps (int16 block offsets)
  std::vector<uint8_t> cl, clm, cues;
  double cl_t0 = 0;
  bool cl_key = false;
  auto flush = [&]() {
    if (cl.empty()) return;
    if (cl_key) {
      // CuePoint { CueTime, CueTrackPositions { CueTrack 1, CueClusterPosition } }
      std::vector<uint8_t> cp, ctp;
      ebml_uint(cp, 0xB3, (uint64_t)(cl_t0 + 0.5));
      ebml_uint(ctp, 0xF7, 1); ebml_uint(ctp, 0xF1, seg_bytes);
      ebml_master(cp, 0xB7, ctp);
      ebml_master(cues, 0xBB, cp);
    }
    clm.clear();
    ebml_master(clm, 0x1F43B675, cl);
    ok = ok && put(f, clm);
    seg_bytes += clm.size();
    cl.clear();
  };
  int64_t k = 0;
  for (; have && ok; have = rd.next(pk), k++) {
    const double ts = frame_ms * k;
    if (cl.empty() || pk.key || ts - cl_t0 > 30000) { flush(); cl_t0 = ts; cl_key = pk.key || k == 0; ebml_uint(cl, 0xE7, (uint64_t)(ts + 0.5)); }
    size_t n = pk.data.size();
    const uint8_t* p = skip_td(pk.data.data(), n);
    const int rel = (int)(ts - cl_t0 + 0.5);
    ebml_id(cl, 0xA3);
    ebml_size(cl, 4 + n);
    cl.push_back(0x81);
    cl.push_back((uint8_t)(rel >> 8)); cl.push_back((uint8_t)rel);
    cl.push_back(pk.key ? 0x80 : 0x00);
    cl.insert(cl.end(), p, p + n);
  }
  flush();
  *n_out = k;
  const uint64_t cues_seg_pos = seg_bytes;
  if (!cues.empty()) {
    clm.clear();
    ebml_master(clm, 0x1C53BB6B, cues);
    ok = ok && put(f, clm);
    seg_bytes += clm.size();
  }
  // patch the Segment size, the Duration and the SeekHead
  {
    const std::vector<uint8_t> shd = mkv_seekhead(info_seg_pos, tracks_seg_pos, cues_seg_pos);
    if (!cues.empty()) ok = ok && fseek(f, seekhead_pos, SEEK_SET) == 0 && put(f, shd);
  }
  std::vector<uint8_t> sz, du;
  ebml_size(sz, seg_bytes);
  ebml_float(du, 0x4489, frame_ms * k);
  ok = ok && fseek(f, seg_size_pos, SEEK_SET) == 0 && put(f, sz);
  ok = ok && fseek(f, info_body_pos + dur_in_info, SEEK_SET) == 0 && fwrite(du.data() + 3, 1, 8, f) == 8;
  return fclose(f) == 0 && ok;
}

bool write_ivf(const std::string& path, PacketReader& rd, int64_t* n_out, int w, int h, int fps_num, int fps_den) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  std::vector<uint8_t> o;
  o.insert(o.end(), {'D', 'K', 'I', 'F'});
  put_le(o, 0, 2); put_le(o, 32, 2);
  o.insert(o.end(), {'A', 'V', '0', '1'});
  put_le(o, (uint64_t)w, 2); put_le(o, (uint64_t)h, 2); put_le(o, (uint64_t)fps_num, 4); put_le(o, (uint64_t)fps_den, 4);
  put_le(o, 0, 4); put_le(o, 0, 4);                              // frame count: patched below
  bool ok = put(f, o);
  Packet pk;
  int64_t k = 0;
  for (; ok && rd.next(pk); k++) {
    o.clear();
    put_le(o, pk.data.size(), 4); put_le(o, (uint64_t)k, 8);
    ok = put(f, o) && put(f, pk.data);
  }
  *n_out = k;
  o.clear(); put_le(o, (uint64_t)k, 4);
  ok = ok && fseek(f, 24, SEEK_SET) == 0 && put(f, o);
  return fclose(f) == 0 && ok;
}

bool write_obu(const std::string& path, PacketReader& rd, int64_t* n_out) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  bool ok = true;
  Packet pk;
  int64_t k = 0;
  for (; ok && rd.next(pk); k++) ok = put(f, pk.data);
  *n_out = k;
  return fclose(f) == 0 && ok;
}

// ---------------------------------------------------------------------------------------------
// GPU leases: concurrent jobs of the daemon (max_concurrent_jobs > 1) must not share a device
// ---------------------------------------------------------------------------------------------
// Never blocks: a job takes the devices that are free now and, when there is none, polls until ANY one is released
// (blocking on a particular device while holding others deadlocks two jobs that each hold what the other waits for).
int lease_device(int dev) {
  char path[128];
  snprintf(path, sizeof(path), "/tmp/av1b200-gpu%d.lock", dev);
  int fd = open(path, O_CREAT | O_RDWR, 0666);
  if (fd < 0) fd = open(path, O_RDONLY);      // another user's lock file: flock needs no write permission
  if (fd < 0) return -1;
  fchmod(fd, 0666);                           // whatever the umask was: the next job may run as another user
  if (flock(fd, LOCK_EX | LOCK_NB) != 0) { close(fd); return -1; }
  return fd;
}

// A group buffer holds kGroup consecutive frames (each Y U V contiguous); the parts cut out of it share it
// and it goes back to the pool when the last of them has been encoded.
// Frames of one read group.  Page-locked when the library can provide it (the encoder's copy engine then
// reads the frames where pread() put them); plain memory otherwise.
struct GroupBuf {
  uint16_t* p = nullptr;
  bool pinned = false;
  GroupBuf(int device, size_t samples) {
    p = static_cast<uint16_t*>(av1b_host_alloc(device, samples * 2));
    pinned = p != nullptr;
    if (!p) p = static_cast<uint16_t*>(malloc(samples * 2));
    if (!p) throw std::bad_alloc();
  }
  ~GroupBuf() { if (pinned) av1b_host_free(p); else free(p); }
  GroupBuf(const GroupBuf&) = delete;
  GroupBuf& operator=(const GroupBuf&) = delete;
  uint16_t* data() { return p; }
};
struct Part {
  int64_t chunk = 0, first_frame = 0;
  bool first_part = false;
  int n = 0;
  std::shared_ptr<GroupBuf> buf;
  size_t first_slot = 0;           // index of the part's first frame inside buf
  // seekable input: a whole chunk as a frame range, the worker reads its frames itself (job_n > 0, no buffer attached)
  int64_t job_first = 0;
  int job_n = 0;
};

struct Shared {
  std::mutex m;
  std::string pkt_dir;                          // finished chunks wait here as packet files
  std::atomic<int64_t> n_chunks{0};
  std::atomic<int64_t> packets{0}, bytes_out{0};
  double fps_num = 30, fps_den = 1;
  // PSNR / SSIM of the frames coded so far: per worker (its encoder's running means of the current chunk + closed chunks)
  std::vector<double> q_psnr_sum, q_ssim_sum, q_frames;
  std::atomic<int64_t> frames_done{0};
  std::atomic<int> failed{0};
  std::string error;
  int64_t total_frames = -1;
  std::chrono::steady_clock::time_point t0;
  std::string progress_path;
  bool quiet = false;
  std::vector<GroupBuf*> free_bufs;   // recycled group buffers (allocation + first touch of ~0.4 GB is slow)
};

std::string chunk_file(const Shared& sh, int64_t chunk) {
  char name[64];
  snprintf(name, sizeof(name), "/chunk_%06lld.pkt", (long long)chunk);
  return sh.pkt_dir + name;
}

// one per worker: the chunk it is coding and that chunk's packet file
struct PacketCtx { Shared* sh; int64_t chunk = -1; FILE* f = nullptr; bool io_error = false; };

int on_packet(void* user, const uint8_t* data, size_t size, int64_t, int is_key) {
  PacketCtx* c = static_cast<PacketCtx*>(user);
  if (!c->f) {
    c->f = fopen(chunk_file(*c->sh, c->chunk).c_str(), "wb");
    if (!c->f) { c->io_error = true; return 1; }
  }
  const uint8_t h[5] = {(uint8_t)size, (uint8_t)(size >> 8), (uint8_t)(size >> 16), (uint8_t)(size >> 24), (uint8_t)(is_key != 0)};
  if (fwrite(h, 1, 5, c->f) != 5 || fwrite(data, 1, size, c->f) != size) { c->io_error = true; return 1; }
  c->sh->packets++; c->sh->bytes_out += (int64_t)size;
  return 0;
}

void report_progress(Shared& sh, bool final_line) {
  const int64_t done = sh.frames_done.load();
  const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - sh.t0).count();
  const double fps = el > 0 ? done / el : 0;
  // the fields of JobMetrics the reference leaves at zero (metrics.rs:12-30, job_executor.rs:117-137)
  const int64_t pk = sh.packets.load();
  const double kbps = pk > 0 ? (double)sh.bytes_out.load() * 8.0 / 1000.0 * (sh.fps_num / sh.fps_den) / (double)pk : 0.0;
  const double eta = (fps > 0 && sh.total_frames > done) ? (double)(sh.total_frames - done) / fps : 0.0;
  double ps = 0, ss = 0, qf = 0;
  {
    std::lock_guard<std::mutex> l(sh.m);
    for (size_t k = 0; k < sh.q_frames.size(); k++) { ps += sh.q_psnr_sum[k]; ss += sh.q_ssim_sum[k]; qf += sh.q_frames[k]; }
  }
  char line[512];
  snprintf(line, sizeof(line), "{\"frames_encoded\": %lld, \"total_frames\": %lld, \"fps\": %.2f, \"bitrate_kbps\": %.1f, "
           "\"est_remaining_secs\": %.1f, \"psnr\": %.3f, \"ssim\": %.5f, \"progress\": %.4f, \"done\": %s}",
           (long long)done, (long long)sh.total_frames, fps, kbps, eta, qf > 0 ? ps / qf : 0.0, qf > 0 ? ss / qf : 0.0,
           sh.total_frames > 0 ? (double)done / (double)sh.total_frames : 0.0, final_line ? "true" : "false");
  if (!sh.quiet) { printf("%s\n", line); fflush(stdout); }
  if (!sh.progress_path.empty()) {
    const std::string tmp = sh.progress_path + ".tmp";
    if (FILE* f = fopen(tmp.c_str(), "w")) { fprintf(f, "%s\n", line); fclose(f); rename(tmp.c_str(), sh.progress_path.c_str()); }
  }
}

}  // namespace

int main(int argc, char** argv) {
  Options o;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    auto val = [&](const char* name) -> std::string {
      if (i + 1 >= argc) die(2, "missing value for %s", name);
      return argv[++i];
    };
    if (a == "--version" || a == "-V") {
      char buf[128];
      av1b_version(buf, sizeof(buf));
      printf("av1an-compatible front end: %s\n", buf);
      return 0;
    } else if (a == "--gpu-plan") {
      // What replaces the daemon's core-count heuristics (concurrency.rs:67-84: 8 workers from 32 cores, 1 job from 24): the
      // unit of parallelism is a GPU.  A job may ask for every GPU (--workers = gpus: the executable leases the free ones and
      // takes at most one per 1200 frames of input), and as many jobs as there are GPUs may run side by side (a queue of short
      // files is start-up bound and runs best as one GPU per job, profiles/r02p_c5_queue_16files_8jobs.json); every job's
      // entropy coder gets cores / workers host threads, and below 4 per GPU the range coder moves onto the device.
      const int gpus = av1b_device_count();
      const int cores = (int)std::max(1u, std::thread::hardware_concurrency());
      if (gpus <= 0) die(4, "no CUDA device visible: this backend has no CPU fallback");
      printf("{\"gpus\": %d, \"host_cores\": %d, \"av1an_workers\": %d, \"max_concurrent_jobs\": %d, \"host_threads_per_gpu\": %d, "
             "\"min_frames_per_worker\": 1200}\n", gpus, cores, gpus, gpus, std::max(1, cores / gpus));
      return 0;
    } else if (a == "--help" || a == "-h") {
      printf("usage: av1an -i IN -o OUT [--encoder svt-av1] [--pix-format yuv420p10le] [--video-params \"--crf N --preset N --keyint N ...\"]\n"
             "             [--audio-params S] [--workers N(GPUs)] [--temp DIR] [--quiet]\n"
             "       av1an --gpu-plan    (JSON: GPUs, workers per job and concurrent jobs for the daemon's ConcurrencyPlan)\n");
      return 0;
    } else if (a == "-i") o.input = val("-i");
    else if (a == "-o") o.output = val("-o");
    else if (a == "--encoder" || a == "-e") o.encoder = val("--encoder");
    else if (a == "--pix-format") o.pix_format = val("--pix-format");
    else if (a == "--video-params" || a == "-v") o.video_params = val("--video-params");
    else if (a == "--audio-params" || a == "-a") o.audio_params = val("--audio-params");
    else if (a == "--workers" || a == "-w") o.workers = atoi(val("--workers").c_str());
    else if (a == "--temp") o.temp = val("--temp");
    else if (a == "--quiet" || a == "-q") o.quiet = true;
    else if (a == "--mux-packets") o.mux_packets = val("--mux-packets");
    else if (a == "--mux-format") o.mux_format = val("--mux-format");
    else if (a == "--no-scene-detection") o.no_scene_detection = true;
    else if (a == "--min-scene-len") o.min_scene_len = atoi(val("--min-scene-len").c_str());
    else if (a == "--sc-method" || a == "--split-method") { if (val(a.c_str()) == "none") o.no_scene_detection = true; }
    else die(2, "unknown argument %s", a.c_str());
  }
  if (!o.mux_packets.empty()) {
    // Container writers alone (no GPU, no input): the packet files DIR/chunk_NNNNNN.pkt (what the workers leave under --temp:
    // 4-byte little-endian size, key flag, temporal unit) are concatenated in order into -o by its extension.  The CPU tests
    // mux oracle-coded streams with it and hand the Matroska file to FFmpeg's demuxer (tests/test_mux.py); also a way to
    // finish a job whose packet files survived.
    if (o.output.empty()) die(2, "-o is required");
    int cw = 0, ch = 0, sw = 0, shh = 0, fn = 30, fd = 1, bits = 10;
    if (sscanf(o.mux_format.c_str(), "%dx%d,%dx%d,%d:%d,%d", &cw, &ch, &sw, &shh, &fn, &fd, &bits) != 7 || cw < 16 || ch < 16 || sw < 1 || sw > cw ||
        shh < 1 || shh > ch || fn < 1 || fd < 1)
      die(2, "--mux-format wants CWxCH,SWxSH,NUM:DEN,BITS (coded size, source size, frame rate, bit depth)");
    Shared msh;
    msh.pkt_dir = o.mux_packets;
    PacketReader rd;
    for (int64_t c = 0;; c++) {
      const std::string f = chunk_file(msh, c);
      if (access(f.c_str(), R_OK) != 0) break;
      rd.files.push_back(f);
    }
    if (rd.files.empty()) die(3, "no packet files (chunk_000000.pkt ...) under %s", o.mux_packets.c_str());
    const std::string tmp_out = o.output + ".part";
    int64_t written = 0;
    const size_t dot = o.output.rfind('.');
    const std::string ext = dot == std::string::npos ? "" : o.output.substr(dot);
    bool ok;
    if (ext == ".ivf") ok = write_ivf(tmp_out, rd, &written, cw, ch, fn, fd);
    else if (ext == ".obu") ok = write_obu(tmp_out, rd, &written);
    else ok = write_mkv(tmp_out, rd, &written, cw, ch, sw, shh, fn, fd, bits > 8);
    if (!ok || written == 0 || rename(tmp_out.c_str(), o.output.c_str()) != 0) { unlink(tmp_out.c_str()); die(6, "cannot write %s", o.output.c_str()); }
    if (!o.quiet) printf("{\"muxed_packets\": %lld}\n", (long long)written);
    return 0;
  }
  if (o.input.empty() || o.output.empty()) die(2, "-i and -o are required");
  if (o.encoder != "svt-av1") die(2, "unsupported --encoder %s (the daemon passes svt-av1)", o.encoder.c_str());
  if (o.workers < 1) die(2, "--workers must be >= 1");
  parse_video_params(o.video_params, o);
  if (o.crf < 0 || o.crf > 63) die(2, "--crf out of range 0..63");
  if (o.keyint < 1) o.keyint = 240;
  int out_bits = 10;
  if (o.pix_format == "yuv420p") out_bits = 8;
  else if (o.pix_format != "yuv420p10le") die(2, "unsupported --pix-format %s (yuv420p, yuv420p10le)", o.pix_format.c_str());

  const int ndev = av1b_device_count();
  if (ndev <= 0) die(4, "no CUDA device visible: this backend has no CPU fallback");
  // AV1B_SHARE_GPU=1 (tests): let several workers share one device so that chunk scheduling and
  // concatenation can be exercised on a single-GPU box
  const bool share = getenv("AV1B_SHARE_GPU") && atoi(getenv("AV1B_SHARE_GPU")) != 0;
  int n_workers = share ? o.workers : std::min(o.workers, ndev);

  // ---- input ----
  Y4m in;
  in.f = fopen(o.input.c_str(), "rb");
  if (!in.f) die(3, "cannot open %s: %s", o.input.c_str(), strerror(errno));
  if (!y4m_parse_header(in)) {
    fclose(in.f);
    if (system("ffmpeg -version > /dev/null 2>&1") != 0)
      die(3, "%s is not YUV4MPEG2 and no ffmpeg is on PATH to decode it", o.input.c_str());
    std::string q;
    for (char c : o.input) { if (c == '\'') q += "'\\''"; else q += c; }
    const std::string cmd = "ffmpeg -v error -i '" + q + "' -map 0:v:0 -pix_fmt " + std::string(out_bits == 8 ? "yuv420p" : "yuv420p10le") + " -strict -1 -f yuv4mpegpipe -";
    in.f = popen(cmd.c_str(), "r");
    in.pipe = true;
    if (!in.f || !y4m_parse_header(in)) die(3, "ffmpeg could not decode %s", o.input.c_str());
  } else {
    struct stat st;
    if (fstat(fileno(in.f), &st) == 0 && S_ISREG(st.st_mode))
      in.n_frames = (st.st_size - in.header_len) / (int64_t)(6 + in.frame_bytes);
  }
  if (in.w < 16 || in.h < 16 || in.w > 8192 || in.h > 4352) die(3, "unsupported frame size %dx%d (16..8192 x 16..4352)", in.w, in.h);
  // sizes that are not multiples of 8 (1920x804 and the like): the library codes the picture padded by edge replication and
  // signals the source size as render_size; the Matroska track carries the padding as PixelCrop
  const int coded_w = (in.w + 7) & ~7, coded_h = (in.h + 7) & ~7;
  const int cw2 = (in.w + 1) / 2, ch2 = (in.h + 1) / 2;     // chroma plane of the source (4:2:0)
  const size_t luma_samples = (size_t)in.w * in.h, chroma_samples = (size_t)cw2 * ch2;
  // A worker costs a CUDA context and an encoder (about 3 s on an 8-GPU box, one after the other in the driver) and codes
  // 4K at some 800 frames/s from a Y4M file: below about 1200 frames per worker more GPUs make a job slower, not faster
  // (profiles/r02p_c4_cli_4k10_2400frames.json), and leave fewer for the other jobs of the queue.  AV1B_MIN_FRAMES_PER_WORKER overrides.
  if (!share && in.n_frames > 0) {
    const char* ev = getenv("AV1B_MIN_FRAMES_PER_WORKER");
    const int64_t per = ev ? std::max<int64_t>(1, atoll(ev)) : 1200;
    n_workers = (int)std::max<int64_t>(1, std::min<int64_t>(n_workers, in.n_frames / per));
  }
  if (in.bits > out_bits) die(3, "input is %d-bit but --pix-format asks for %d-bit", in.bits, out_bits);
  const int shift = out_bits - in.bits;
  const size_t frame_samples = luma_samples + 2 * chroma_samples;

  Shared sh;
  sh.total_frames = in.n_frames;
  sh.t0 = std::chrono::steady_clock::now();
  sh.quiet = o.quiet;
  sh.fps_num = in.fps_num; sh.fps_den = in.fps_den;
  if (!o.temp.empty()) { mkdir(o.temp.c_str(), 0777); sh.progress_path = o.temp + "/progress.json"; sh.pkt_dir = o.temp; }
  else { sh.pkt_dir = o.output + ".av1b200.tmp"; mkdir(sh.pkt_dir.c_str(), 0777); }

  // ---- devices ----
  std::vector<int> dev, lease;
  for (int waited_ms = 0; dev.empty(); waited_ms += 200) {
    for (int d = 0; d < ndev && (int)dev.size() < n_workers; d++) {
      const int fd = lease_device(d);
      if (fd >= 0) { dev.push_back(d); lease.push_back(fd); }
    }
    if (!dev.empty()) break;
    if (waited_ms == 0 && !o.quiet) fprintf(stderr, "av1an (av1b200): all %d GPUs are leased by other jobs, waiting for one\n", ndev);
    usleep(200 * 1000);
  }
  if (share) while ((int)dev.size() < n_workers) dev.push_back(dev[dev.size() % lease.size()]);

  // ---- workers: one encoder handle (= one GPU) each, fed with chunk parts through a bounded queue ----
  const int kPart = 16;   // frames per hand-over: two device batches; buffers are recycled through sh.free_bufs
  // AV1B_CLI_TIMING=1: where the wall clock of the job goes (stderr, at the end)
  const bool timing = getenv("AV1B_CLI_TIMING") != nullptr;
  const auto t_job0 = std::chrono::steady_clock::now();
  auto since = [](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
  std::atomic<int64_t> us_create{0}, us_encode{0}, us_flush{0}, us_wread{0};
  double s_alloc = 0, s_read = 0, s_submit = 0;
  // seekable Y4M files: frames [first, first + n) into dst (uint16 planes, one frame after the other) with parallel
  // pread(2), one frame per task; returns 0, 1 (truncated) or 2 (a frame header that is not the plain marker)
  const int in_fd = in.pipe ? -1 : fileno(in.f);
  const int kJobPart = 8;            // frames per hand-over of a self-feeding worker: one device batch
  const int hw_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  const int worker_io_threads = std::max(1, std::min(8, hw_threads / (2 * std::max(1, (int)dev.size()))));
  auto read_frames = [&](int64_t first, int n, uint16_t* dst, int n_threads) -> int {
    std::atomic<int> next{0}, bad{0};
    auto work = [&]() {
      std::vector<uint8_t> tmp;
      for (;;) {
        const int k = next.fetch_add(1);
        if (k >= n || bad) break;
        uint16_t* slot = dst + (size_t)k * frame_samples;
        const off_t off = (off_t)in.header_len + (off_t)(first + k) * (off_t)(6 + in.frame_bytes) + 6;
        char mark[6];
        if (pread(in_fd, mark, 6, off - 6) != 6 || memcmp(mark, "FRAME\n", 6) != 0) { bad = 2; break; }
        uint8_t* d8 = in.bits > 8 ? reinterpret_cast<uint8_t*>(slot) : (tmp.resize(in.frame_bytes), tmp.data());
        size_t done = 0;
        while (done < in.frame_bytes) {
          const ssize_t r = pread(in_fd, d8 + done, in.frame_bytes - done, off + (off_t)done);
          if (r <= 0) { bad = 1; break; }
          done += (size_t)r;
        }
        if (in.bits <= 8) for (size_t i = 0; i < in.frame_bytes; i++) slot[i] = (uint16_t)(tmp[i] << shift);
      }
    };
    std::vector<std::thread> io;
    for (int t = 1; t < n_threads; t++) io.emplace_back(work);
    work();
    for (auto& t : io) t.join();
    return bad.load();
  };
  struct Queue { std::mutex m; std::condition_variable cv; std::deque<Part> q; bool closed = false; };
  const int W = (int)dev.size();
  sh.q_psnr_sum.assign(W, 0); sh.q_ssim_sum.assign(W, 0); sh.q_frames.assign(W, 0);
  std::vector<Queue> queues(W);
  std::vector<std::thread> threads;
  for (int wk = 0; wk < W; wk++) {
    threads.emplace_back([&, wk]() {
      av1b_config cfg;
      av1b_config_default(&cfg);
      cfg.width = in.w; cfg.height = in.h; cfg.bit_depth = out_bits;
      cfg.fps_num = in.fps_num; cfg.fps_den = in.fps_den;
      cfg.crf = o.crf; cfg.preset = o.preset; cfg.keyint = o.keyint; cfg.lookahead = o.lookahead;
      cfg.film_grain = o.film_grain; cfg.enable_qm = o.enable_qm; cfg.qm_min = o.qm_min; cfg.qm_max = o.qm_max;
      cfg.device_id = dev[wk];
      cfg.tune[3] = 1;                 // PSNR / SSIM for the progress events
      cfg.host_threads = std::max(1u, std::thread::hardware_concurrency() / (unsigned)W);
      av1b_encoder* enc = nullptr;
      std::unique_ptr<GroupBuf> own_buf[2];   // seekable input: the part buffers this worker reads into
      PacketCtx ctx;
      ctx.sh = &sh;
      const auto tc0 = std::chrono::steady_clock::now();
      int rc = av1b_encoder_create(&cfg, &enc);
      us_create += (int64_t)(since(tc0) * 1e6);
      if (rc != AV1B_OK) {
        std::lock_guard<std::mutex> l(sh.m);
        if (!sh.failed) { sh.failed = -rc; sh.error = av1b_last_error(); }
      }
      for (;;) {
        Part part;
        {
          std::unique_lock<std::mutex> l(queues[wk].m);
          queues[wk].cv.wait(l, [&] { return !queues[wk].q.empty() || queues[wk].closed; });
          if (queues[wk].q.empty()) break;
          part = std::move(queues[wk].q.front());
          queues[wk].q.pop_front();
        }
        queues[wk].cv.notify_all();
        if (sh.failed || !enc) continue;   // drain
        auto end_chunk = [&]() {
          // end of a chunk: everything still in flight belongs to it; then its packet file is complete
          const auto tf0 = std::chrono::steady_clock::now();
          rc = av1b_encode_flush(enc, on_packet, nullptr, &ctx);
          us_flush += (int64_t)(since(tf0) * 1e6);
          if (ctx.f) { if (fclose(ctx.f) != 0) ctx.io_error = true; ctx.f = nullptr; }
          double ps = 0, ss = 0; int64_t qf = 0;
          if (av1b_get_quality(enc, &ps, &ss, &qf) == AV1B_OK) {
            std::lock_guard<std::mutex> l(sh.m);
            sh.q_psnr_sum[wk] += ps * qf; sh.q_ssim_sum[wk] += ss * qf; sh.q_frames[wk] += (double)qf;
          }
          if (rc != AV1B_OK || ctx.io_error) {
            std::lock_guard<std::mutex> l(sh.m);
            if (!sh.failed) { sh.failed = ctx.io_error ? 6 : -rc; sh.error = ctx.io_error ? "cannot write the chunk's packet file (disk full?)" : av1b_last_error(); }
          }
        };
        if (part.job_n > 0) {
          // a whole chunk of a seekable file: this worker reads its frames itself, a part at a time into its own page-locked
          // buffer (av1b_encode_stream returns once the part has been uploaded, so the next read overlaps the kernels);
          // the workers read side by side, nobody waits for a common reader
          ctx.chunk = part.chunk;
          for (auto& ob : own_buf) if (!ob) ob.reset(new GroupBuf(dev[wk], (size_t)kJobPart * frame_samples));
          // two part buffers: the next part is read (by a helper thread) while the current one is uploaded and coded
          auto read_part = [&](int f0, int which) -> int {
            const int n = std::min(kJobPart, part.job_n - f0);
            const auto tr0 = std::chrono::steady_clock::now();
            const int bad = read_frames(part.job_first + f0, n, own_buf[which]->data(), worker_io_threads);
            us_wread += (int64_t)(since(tr0) * 1e6);
            return bad;
          };
          int bad = read_part(0, 0);
          for (int f0 = 0, which = 0; f0 < part.job_n && !sh.failed; f0 += kJobPart, which ^= 1) {
            const int n = std::min(kJobPart, part.job_n - f0);
            if (bad) {
              std::lock_guard<std::mutex> l(sh.m);
              if (!sh.failed) { sh.failed = 3; sh.error = bad == 2 ? "Y4M frame headers carry parameters (or the file is corrupt): pipe it through ffmpeg or rewrite it with plain FRAME markers" : "truncated Y4M file"; }
              break;
            }
            int next_bad = 0;
            std::thread ahead;
            if (f0 + kJobPart < part.job_n) ahead = std::thread([&, f0, which]() { next_bad = read_part(f0 + kJobPart, which ^ 1); });
            std::vector<av1b_frame_src> fs((size_t)n);
            for (int k = 0; k < n; k++) {
              uint16_t* b = own_buf[which]->data() + (size_t)k * frame_samples;
              fs[k].planes[0] = b; fs[k].planes[1] = b + luma_samples; fs[k].planes[2] = b + luma_samples + chroma_samples;
              fs[k].stride[0] = in.w; fs[k].stride[1] = fs[k].stride[2] = cw2;
            }
            const auto te0 = std::chrono::steady_clock::now();
            rc = av1b_encode_stream(enc, fs.data(), (uint32_t)n, f0 == 0 ? 1 : 0, part.job_first + f0, on_packet, nullptr, &ctx);
            us_encode += (int64_t)(since(te0) * 1e6);
            if (ahead.joinable()) ahead.join();
            bad = next_bad;
            if (rc != AV1B_OK) {
              std::lock_guard<std::mutex> l(sh.m);
              if (!sh.failed) { sh.failed = -rc; sh.error = av1b_last_error(); }
              break;
            }
            sh.frames_done += n;
          }
          if (!sh.failed) end_chunk();
          continue;
        }
        if (part.n == 0) { end_chunk(); continue; }
        ctx.chunk = part.chunk;
        std::vector<av1b_frame_src> fs((size_t)part.n);
        for (int k = 0; k < part.n; k++) {
          uint16_t* b = part.buf->data() + (part.first_slot + (size_t)k) * frame_samples;
          fs[k].planes[0] = b; fs[k].planes[1] = b + luma_samples; fs[k].planes[2] = b + luma_samples + chroma_samples;
          fs[k].stride[0] = in.w; fs[k].stride[1] = fs[k].stride[2] = cw2;
        }
        const auto te0 = std::chrono::steady_clock::now();
        rc = av1b_encode_stream(enc, fs.data(), (uint32_t)part.n, part.first_part ? 1 : 0, part.first_frame, on_packet, nullptr, &ctx);
        us_encode += (int64_t)(since(te0) * 1e6);
        part.buf.reset();   // the last part of a group returns the buffer to the pool (custom deleter)
        if (rc != AV1B_OK) {
          std::lock_guard<std::mutex> l(sh.m);
          if (!sh.failed) { sh.failed = -rc; sh.error = av1b_last_error(); }
          continue;
        }
        sh.frames_done += part.n;
      }
      if (enc) av1b_encoder_destroy(enc);
    });
  }

  // ---- reader: av1an-style chunking = a new closed GOP at every detected scene cut and at the latest after
  //      --keyint frames; chunk c -> worker c mod W.  A pipe is read sequentially in groups of kPart frames that travel to
  //      the workers as parts; of a regular Y4M file the reader only samples the thumbnail rows, the workers read the
  //      frames of their chunks themselves with parallel pread(2) (one frame per task). ----
  auto take_buffer = [&]() -> std::shared_ptr<GroupBuf> {
    GroupBuf* g = nullptr;
    {
      std::lock_guard<std::mutex> l(sh.m);
      if (!sh.free_bufs.empty()) { g = sh.free_bufs.back(); sh.free_bufs.pop_back(); }
    }
    if (!g) g = new GroupBuf(dev[0], (size_t)kPart * frame_samples);
    Shared* shp = &sh;
    return std::shared_ptr<GroupBuf>(g, [shp](GroupBuf* p) { std::lock_guard<std::mutex> l(shp->m); shp->free_bufs.push_back(p); });
  };
  auto submit = [&](Part&& part) {
    if (part.chunk + 1 > sh.n_chunks) sh.n_chunks = part.chunk + 1;
    Queue& q = queues[(size_t)(part.chunk % W)];
    const auto ts0 = std::chrono::steady_clock::now();
    {
      std::unique_lock<std::mutex> l(q.m);
      q.cv.wait(l, [&] { return q.q.size() < 3; });
      q.q.push_back(std::move(part));
    }
    q.cv.notify_all();
    s_submit += since(ts0);
  };
  const int tw = (in.w - 4 + 7) / 8, th = (in.h - 4 + 7) / 8;      // thumbnail: every 8th sample from (4, 4)
  auto make_thumb = [&](const uint16_t* luma, uint16_t* t) {
    for (int y = 4, k = 0; y < in.h; y += 8) for (int x = 4; x < in.w; x += 8) t[k++] = luma[(size_t)y * in.w + x];
  };
  const int fd = in.pipe ? -1 : fileno(in.f);
  const int io_threads = in.pipe ? 1 : (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
  std::vector<uint16_t> thumbs((size_t)kPart * tw * th), thumb_prev;
  std::vector<uint8_t> raw;
  int64_t frame = 0, chunk = 0;
  int in_chunk = 0;
  double score_avg = -1;
  bool eof = false;
  auto last_report = std::chrono::steady_clock::now();
  // scene-cut score: mean absolute luma difference on the 1/8 x 1/8 thumbnails, in 8-bit units;
  // a cut = a jump well above both an absolute floor and the recent level of change
  auto is_cut = [&](const uint16_t* tc) -> bool {
    bool cut = false;
    if (!thumb_prev.empty() && !o.no_scene_detection) {
      uint64_t sad = 0;
      for (int i = 0; i < tw * th; i++) sad += (uint64_t)std::abs((int)tc[i] - (int)thumb_prev[i]);
      const double score = (double)sad / (tw * th) / (1 << (out_bits - 8));
      if (in_chunk >= o.min_scene_len && score > 10.0 && (score_avg < 0 || score > 3.0 * score_avg + 2.0)) cut = true;
      score_avg = score_avg < 0 ? score : 0.8 * score_avg + 0.2 * score;
      if (cut) score_avg = -1;
    }
    thumb_prev.assign(tc, tc + (size_t)tw * th);
    return cut;
  };
  const bool seekable = !in.pipe && in.n_frames >= 0;
  if (seekable) {
    // A regular Y4M file: the reader only looks at the thumbnails (the sampled rows: a twelfth of the file), decides the
    // cuts and hands whole chunks to the workers as frame ranges; every worker reads its own frames (see above).
    const int kScan = 64;
    const size_t bps = in.bits > 8 ? 2 : 1;
    std::vector<uint16_t> sthumbs((size_t)kScan * tw * th);
    int64_t chunk_first = 0;
    auto submit_job = [&](int64_t first, int64_t n) {
      if (n <= 0) return;
      Part job; job.chunk = chunk; job.job_first = first; job.job_n = (int)n;
      submit(std::move(job));
    };
    while (frame < in.n_frames && !sh.failed) {
      const auto tr0 = std::chrono::steady_clock::now();
      const int got = (int)std::min<int64_t>(kScan, in.n_frames - frame);
      std::atomic<int> next{0}, bad{0};
      auto work = [&]() {
        std::vector<uint8_t> row((size_t)in.w * bps);
        for (;;) {
          const int k = next.fetch_add(1);
          if (k >= got || bad) break;
          const off_t off = (off_t)in.header_len + (off_t)(frame + k) * (off_t)(6 + in.frame_bytes) + 6;
          char mark[6];
          if (pread(fd, mark, 6, off - 6) != 6 || memcmp(mark, "FRAME\n", 6) != 0) { bad = 2; break; }
          uint16_t* t = sthumbs.data() + (size_t)k * tw * th;
          int q = 0;
          for (int y = 4; y < in.h && !bad; y += 8) {
            if (pread(fd, row.data(), row.size(), off + (off_t)y * (off_t)row.size()) != (ssize_t)row.size()) { bad = 1; break; }
            if (bps == 2) { const uint16_t* r16 = reinterpret_cast<const uint16_t*>(row.data()); for (int x = 4; x < in.w; x += 8) t[q++] = r16[x]; }
            else for (int x = 4; x < in.w; x += 8) t[q++] = (uint16_t)(row[x] << shift);
          }
        }
      };
      std::vector<std::thread> io;
      for (int t = 1; t < io_threads; t++) io.emplace_back(work);
      work();
      for (auto& t : io) t.join();
      s_read += since(tr0);
      if (bad == 2) die(3, "Y4M frame headers carry parameters (or the file is corrupt): pipe it through ffmpeg or rewrite it with plain FRAME markers");
      if (bad) die(3, "truncated Y4M file");
      for (int k = 0; k < got; k++) {
        const bool cut = is_cut(sthumbs.data() + (size_t)k * tw * th);
        if (frame > 0 && (cut || in_chunk >= o.keyint)) {
          submit_job(chunk_first, frame - chunk_first);
          chunk_first = frame; chunk++; in_chunk = 0;
        }
        frame++; in_chunk++;
      }
      const auto now = std::chrono::steady_clock::now();
      if (std::chrono::duration<double>(now - last_report).count() > 1.0) { report_progress(sh, false); last_report = now; }
    }
    submit_job(chunk_first, frame - chunk_first);
    eof = true;
  }
  while (!eof && !sh.failed) {
    const auto ta0 = std::chrono::steady_clock::now();
    std::shared_ptr<GroupBuf> buf = take_buffer();
    s_alloc += since(ta0);
    const auto tr0 = std::chrono::steady_clock::now();
    int got = 0;
    while (got < kPart) {   // a pipe (or a file of unknown length): frames arrive in order
      uint16_t* slot = buf->data() + (size_t)got * frame_samples;
      if (!y4m_read_frame(in, raw, slot, shift)) { eof = true; break; }
      make_thumb(slot, thumbs.data() + (size_t)got * tw * th);
      got++;
    }
    s_read += since(tr0);
    // cut decisions in display order; a part never crosses a chunk boundary
    int run_start = 0;
    auto emit = [&](int from, int to) {
      if (to <= from) return;
      Part part;
      part.chunk = chunk; part.first_frame = frame - (to - from); part.first_part = (in_chunk - (to - from)) == 0;
      part.n = to - from; part.buf = buf; part.first_slot = (size_t)from;
      submit(std::move(part));
    };
    for (int k = 0; k < got; k++) {
      const bool cut = is_cut(thumbs.data() + (size_t)k * tw * th);
      if (frame > 0 && (cut || in_chunk >= o.keyint)) {
        emit(run_start, k);
        run_start = k;
        { Part end; end.chunk = chunk; end.n = 0; submit(std::move(end)); }   // end of chunk: flush
        chunk++; in_chunk = 0;
      }
      frame++; in_chunk++;
    }
    emit(run_start, got);
    const auto now = std::chrono::steady_clock::now();
    if (std::chrono::duration<double>(now - last_report).count() > 1.0) { report_progress(sh, false); last_report = now; }
  }
  if (frame > 0 && !seekable) { Part end; end.chunk = chunk; end.n = 0; submit(std::move(end)); }
  for (auto& q : queues) { { std::lock_guard<std::mutex> l(q.m); q.closed = true; } q.cv.notify_all(); }
  // the workers of a seekable job still have whole chunks ahead of them: keep the progress events coming
  {
    std::atomic<bool> joined{false};
    std::thread reporter([&]() {
      auto last = std::chrono::steady_clock::now();
      while (!joined) {
        usleep(100 * 1000);
        const auto now = std::chrono::steady_clock::now();
        if (!joined && std::chrono::duration<double>(now - last).count() > 1.0) { report_progress(sh, false); last = now; }
      }
    });
    for (auto& t : threads) t.join();
    joined = true;
    reporter.join();
  }
  const double s_encode_phase = since(t_job0);
  auto cleanup_packets = [&]() {
    for (int64_t c = 0; c < sh.n_chunks; c++) unlink(chunk_file(sh, c).c_str());
    if (o.temp.empty()) rmdir(sh.pkt_dir.c_str());
  };
  if (in.pipe) {
    // a decoder that died looks like end of stream to the reader: its exit status is what tells a truncated job from a whole one
    const int st = pclose(in.f);
    if (!(st != -1 && WIFEXITED(st) && WEXITSTATUS(st) == 0)) {
      cleanup_packets(); unlink(o.output.c_str());
      die(3, "the ffmpeg decode pipe failed (status %d): the input was not read completely", st);
    }
  } else fclose(in.f);
  for (int lfd : lease) close(lfd);
  for (GroupBuf* gb : sh.free_bufs) delete gb;
  sh.free_bufs.clear();
  if (sh.total_frames < 0) sh.total_frames = frame;

  if (sh.failed) {
    cleanup_packets();
    unlink(o.output.c_str());
    die(sh.failed, "encode failed: %s", sh.error.c_str());
  }
  if (frame == 0) { cleanup_packets(); die(3, "input has no frames"); }

  // ---- concatenate the chunk streams in order and write the container (streamed from the packet files) ----
  PacketReader rd;
  for (int64_t c = 0; c < sh.n_chunks; c++) rd.files.push_back(chunk_file(sh, c));
  const std::string tmp_out = o.output + ".part";
  bool ok;
  int64_t written = 0;
  const size_t dot = o.output.rfind('.');
  const std::string ext = dot == std::string::npos ? "" : o.output.substr(dot);
  if (ext == ".ivf") ok = write_ivf(tmp_out, rd, &written, coded_w, coded_h, in.fps_num, in.fps_den);
  else if (ext == ".obu") ok = write_obu(tmp_out, rd, &written);
  else ok = write_mkv(tmp_out, rd, &written, coded_w, coded_h, in.w, in.h, in.fps_num, in.fps_den, out_bits > 8);
  cleanup_packets();
  if (ok && written != frame) { unlink(tmp_out.c_str()); unlink(o.output.c_str()); die(5, "internal error: %lld packets for %lld frames", (long long)written, (long long)frame); }
  if (ok && in.pipe && ext != ".ivf" && ext != ".obu" && o.audio_params.find("copy") != std::string::npos) {
    // the source went through ffmpeg, so ffmpeg is available: copy its audio streams next to our video.  The daemon asked
    // for the audio (av1an.rs:97) and replaces the source with our output: a failed copy is a failed job, not a silent loss
    std::string qi, qo;
    for (char c : o.input) { if (c == '\'') qi += "'\\''"; else qi += c; }
    for (char c : tmp_out) { if (c == '\'') qo += "'\\''"; else qo += c; }
    const std::string mux = "ffmpeg -v error -y -i '" + qo + "' -i '" + qi + "' -map 0:v:0 -map 1:a? -c copy -f matroska '" + qo + ".mux'";
    const int st = system(mux.c_str());
    if (st != -1 && WIFEXITED(st) && WEXITSTATUS(st) == 0 && rename((tmp_out + ".mux").c_str(), tmp_out.c_str()) == 0) {
    } else {
      unlink((tmp_out + ".mux").c_str()); unlink(tmp_out.c_str()); unlink(o.output.c_str());
      die(7, "copying the audio streams failed (ffmpeg status %d): no output written", st);
    }
  }
  if (!ok || rename(tmp_out.c_str(), o.output.c_str()) != 0) { unlink(tmp_out.c_str()); die(6, "cannot write %s", o.output.c_str()); }
  report_progress(sh, true);
  if (timing)
    fprintf(stderr, "av1an timing: total %.2f s = read+encode phase %.2f s + container %.2f s; reader: buffers %.2f s, read %.2f s, waiting for workers %.2f s; "
                    "workers (summed over %d): encoder create %.2f s, own reads %.2f s, encode calls %.2f s, flushes %.2f s\n",
            since(t_job0), s_encode_phase, since(t_job0) - s_encode_phase, s_alloc, s_read, s_submit, W, us_create / 1e6, us_wread / 1e6, us_encode / 1e6, us_flush / 1e6);
  return 0;
}
