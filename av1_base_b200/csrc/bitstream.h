// Host side of the encode path: AV1 OBU packing and tile entropy coding (AV1 spec sections 5, 8).
// "Tile/superblock entropy coding runs on the host over the device-produced symbol streams"
// (BASELINE.json north_star); replaces what av1an + SVT-AV1 do behind
// /root/reference/crates/daemon/src/encode/av1an.rs:126-139 (run_av1an).
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <emmintrin.h>
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
// Everything on the per-symbol path is inline: the tile writers call it a few hundred thousand
// times per frame.
class RangeEncoder {
 public:
  explicit RangeEncoder(bool adapt) : adapt_(adapt) { grow(1 << 14); }
  ~RangeEncoder() { free(pre_); }
  RangeEncoder(const RangeEncoder&) = delete;
  RangeEncoder& operator=(const RangeEncoder&) = delete;
  // icdf: inverted CDF (32768 - cdf), n symbols, icdf[n-1] == 0, icdf[n] = adaptation counter.
  // The adaptation (spec 8.2.6, expressed on the inverted CDF) is written without data-dependent branches:
  // the symbol value and the alphabet size change from call to call, a loop over n - 1 entries mispredicts
  // its exit about once a symbol, and so does a switch on the alphabet size: every alphabet takes the same SSE2 path.
  __attribute__((always_inline)) inline void symbol(int s, uint16_t* icdf, int n) {
    encode(s, icdf, n);
    if (!adapt_) return;
    const int cnt = icdf[n];
    const int rate = 3 + (cnt > 15) + (cnt > 31) + (n > 3 ? 2 : 1);   // + min(floor(log2(n)), 2)
    {
      // one SSE2 step updates entries 0..7 (all of an alphabet of up to 9 symbols), a second one entries 8..15:
      // lanes i < s move towards 32768, lanes s <= i < n - 1 towards 0, lanes >= n - 1 keep their value (they belong
      // to the counter / the next CDF: every CDF handed to symbol() has 16 readable and writable entries behind its start)
      const __m128i sh = _mm_cvtsi32_si128(rate), top = _mm_set1_epi16((short)0x8000);
      const __m128i vs = _mm_set1_epi16((short)s), vn = _mm_set1_epi16((short)(n - 1));
      __m128i i8 = _mm_setr_epi16(0, 1, 2, 3, 4, 5, 6, 7);
      for (int k = 0; k < n - 1; k += 8) {
        const __m128i x = _mm_loadu_si128(reinterpret_cast<const __m128i*>(icdf + k));
        const __m128i up = _mm_add_epi16(x, _mm_srl_epi16(_mm_sub_epi16(top, x), sh));
        const __m128i dn = _mm_sub_epi16(x, _mm_srl_epi16(x, sh));
        const __m128i m_up = _mm_cmplt_epi16(i8, vs), m_live = _mm_cmplt_epi16(i8, vn);
        __m128i y = _mm_or_si128(_mm_and_si128(m_up, up), _mm_andnot_si128(m_up, dn));
        y = _mm_or_si128(_mm_and_si128(m_live, y), _mm_andnot_si128(m_live, x));
        _mm_storeu_si128(reinterpret_cast<__m128i*>(icdf + k), y);
        i8 = _mm_add_epi16(i8, _mm_set1_epi16(8));
      }
    }
    icdf[n] = (uint16_t)(cnt + (cnt < 32));
  }
  // binary symbol with an adaptive CDF: icdf[0] = 32768 - P(0), icdf[1] = 0, icdf[2] = counter
  inline void bit(int b, uint16_t* icdf) {
    encode(b, icdf, 2);
    if (adapt_) {
      const int cnt = icdf[2];
      const int rate = 4 + (cnt > 15) + (cnt > 31);
      if (b) icdf[0] += (uint16_t)((32768 - icdf[0]) >> rate);
      else icdf[0] -= (uint16_t)(icdf[0] >> rate);
      icdf[2] = (uint16_t)(cnt + (cnt < 32));
    }
  }
  inline void boolean(int b) { const uint16_t c[3] = {16384, 0, 0}; encode(b, c, 2); }
  inline void literal(uint32_t v, int n) { for (int i = n - 1; i >= 0; i--) boolean((v >> i) & 1); }
  // terminates the stream and appends the bytes to out
  void finish(std::vector<uint8_t>& out);
 private:
  inline void put(uint16_t v) { if (n_ == cap_) grow(cap_ * 2); pre_[n_++] = v; }
  void grow(size_t cap) {
    pre_ = static_cast<uint16_t*>(realloc(pre_, cap * sizeof(uint16_t)));
    cap_ = cap;
  }
  __attribute__((always_inline)) inline void encode(int s, const uint16_t* icdf, int n) {
    // The decoder partitions [0, rng) from the top: symbol k owns [cur_k, cur_{k-1}) with
    // cur_k = ((rng >> 8) * (icdf[k] >> 6) >> 1) + 4 * (n - 1 - k), cur_{-1} = rng.
    uint32_t r = rng_, l = low_;
    const int N = n - 1;
    const uint32_t v = ((r >> 8) * (uint32_t)(icdf[s] >> 6) >> 1) + 4 * (N - s);
    // symbol 0 owns the top of the range (u = r); a select instead of a branch on the symbol value
    const int nz = s > 0;
    const uint32_t uc = ((r >> 8) * (uint32_t)(icdf[s - nz] >> 6) >> 1) + 4 * (N - s + 1);
    const uint32_t u = nz ? uc : r;
    l += r - u;
    r = u - v;
    const int d = __builtin_clz(r) - 16;   // r < 2^16: make bit 15 the top bit
    int c = cnt_;
    int sft = c + d;
    if (sft >= 0) {
      c += 16;
      uint32_t m = (1u << c) - 1;
      if (sft >= 8) {
        put((uint16_t)(l >> c));
        l &= m;
        c -= 8;
        m >>= 8;
      }
      put((uint16_t)(l >> c));
      sft = c + d - 24;
      l &= m;
    }
    low_ = l << d;
    rng_ = r << d;
    cnt_ = sft;
  }
  bool adapt_;
  uint16_t* pre_ = nullptr;
  size_t n_ = 0, cap_ = 0;
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
// Tile payload of an inter frame from its token list (tokens.h): range coding + CDF adaptation only.
void pack_tile_tokens(const Av1bFrameParams& fp, const uint32_t* tok, size_t n, std::vector<uint8_t>& out);
// default CDF set of a tile (TileCdfs image, tile_cdfs_size() bytes) for the device range coder (rc_kernel.cu)
void tile_cdfs_default(int base_q_idx, void* dst);
size_t tile_cdfs_size();
// CPU statement of the device tokenizer (token_kernel.cu) for an inter frame: digests the raster levels like
// the inter kernel does, derives the mode classes and walks the blocks in coding order; one token list per
// tile.  Test infrastructure for the token path (the product tokenizes on the device).
void tokenize_frame_host(const Av1bFrameParams& fp, const Av1bSeqParams& seq, const Av1bGeom& g, const Av1bFrameSyms& syms,
                         std::vector<std::vector<uint32_t>>& tiles);

// One OBU_FRAME (frame header + all tiles). n_threads > 1 entropy-codes tiles in parallel.
// Returns 0 on success.
int write_frame(const Av1bSeqParams& seq, const Av1bFrameParams& fp, const Av1bGeom& g,
                const Av1bFrameSyms& syms, std::vector<uint8_t>& out, int n_threads);

}  // namespace av1b
