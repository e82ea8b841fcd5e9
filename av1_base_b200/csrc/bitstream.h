// Host side of the encode path: AV1 OBU packing and tile entropy coding (AV1 spec sections 5, 8).
// "Tile/superblock entropy coding runs on the host over the device-produced symbol streams"
// (BASELINE.json north_star); replaces what av1an + SVT-AV1 do behind
// /root/reference/crates/daemon/src/encode/av1an.rs:126-139 (run_av1an).
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include "av1b_types.h"

namespace av1b {

// MSB-first bit writer for uncompressed headers.
class BitWriter {
 public:
  void put(uint32_t v, int n) { for (int i = n - 1; i >= 0; i--) bit((v >> i) & 1); }
  void bit(int b) {
    if (nbits_ == 0) buf_.push_back(0);
    if (b) buf_.back() |= (uint8_t)(0x80 >> nbits_);
    nbits_ = (nbits_ + 1) & 7;
  }
  void byte_align() { while (nbits_) bit(0); }
  void trailing_bits() { bit(1); byte_align(); }
  std::vector<uint8_t>& bytes() { return buf_; }
 private:
  std::vector<uint8_t> buf_;
  int nbits_ = 0;
};

// Multi-symbol range encoder producing the stream the AV1 symbol decoder (spec 8.2) reads.
class RangeEncoder {
 public:
  explicit RangeEncoder(bool adapt) : adapt_(adapt) { pre_.reserve(1 << 16); }
  // icdf: inverted CDF (32768 - cdf), n symbols, icdf[n-1] == 0, icdf[n] = adaptation counter
  void symbol(int s, uint16_t* icdf, int n);
  void boolean(int b) { uint16_t c[3] = {16384, 0, 0}; encode(b, c, 2); }
  void literal(uint32_t v, int n) { for (int i = n - 1; i >= 0; i--) boolean((v >> i) & 1); }
  // terminates the stream and appends the bytes to out
  void finish(std::vector<uint8_t>& out);
 private:
  void encode(int s, const uint16_t* icdf, int n);
  bool adapt_;
  std::vector<uint16_t> pre_;
  uint32_t low_ = 0;
  uint32_t rng_ = 0x8000;
  int cnt_ = -9;
};

void append_obu(std::vector<uint8_t>& out, int obu_type, const std::vector<uint8_t>& payload);
void write_temporal_delimiter(std::vector<uint8_t>& out);
void write_sequence_header(const Av1bSeqParams& seq, std::vector<uint8_t>& out);
// Frame packing split into independent per-tile tasks (tiles share no entropy state), so that the
// encoder can spread (frame, tile) tasks of a whole batch over its host thread pool.
struct FramePack {
  std::vector<uint8_t> header;                  // frame header + tile group header, byte aligned
  std::vector<std::vector<uint8_t>> tiles;      // entropy-coded tile payloads
};
void pack_frame_header(const Av1bSeqParams& seq, const Av1bFrameParams& fp, const Av1bGeom& g, FramePack& fpk);
void pack_tile(const Av1bSeqParams& seq, const Av1bFrameParams& fp, const Av1bGeom& g, const Av1bFrameSyms& syms,
               int tile, std::vector<uint8_t>& out);
void assemble_frame(const FramePack& fpk, std::vector<uint8_t>& out);   // appends one OBU_FRAME

// One OBU_FRAME (frame header + all tiles). n_threads > 1 entropy-codes tiles in parallel.
// Returns 0 on success.
int write_frame(const Av1bSeqParams& seq, const Av1bFrameParams& fp, const Av1bGeom& g,
                const Av1bFrameSyms& syms, std::vector<uint8_t>& out, int n_threads);

}  // namespace av1b
