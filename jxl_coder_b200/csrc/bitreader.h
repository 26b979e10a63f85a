// LSB-first bit reader over a 4-byte-aligned, zero-padded byte buffer (host or HBM).
// JPEG XL packs fields least-significant-bit first (ISO/IEC 18181-1 §B; SURVEY.md App. B notation u(n)).
#pragma once
#include "hd.h"

namespace jxlb {

struct BitReader {
  const uint32_t* words;  // 4-byte aligned base of the whole buffer
  uint64_t buf;           // unread bits, next bit at bit 0
  uint32_t nbits;         // number of valid bits in buf
  uint32_t widx;          // next word to fetch
  uint32_t wend;          // words readable (reads past it yield zeros)
  uint64_t start_bit;     // absolute bit position of the first bit of this stream
  uint64_t end_bit;       // absolute bit position one past the stream's last bit (for the overrun check)

  // base must be 4-byte aligned; buffer_bytes is the readable size of the buffer (multiple of 4 after padding).
  JXLB_HD void Init(const uint8_t* base, uint64_t buffer_bytes, uint64_t start_bit_, uint64_t end_bit_) {
    words = reinterpret_cast<const uint32_t*>(base);
    wend = (uint32_t) (buffer_bytes >> 2);
    start_bit = start_bit_;
    end_bit = end_bit_;
    widx = (uint32_t) (start_bit_ >> 5);
    buf = 0;
    nbits = 0;
    Refill();
    uint32_t skip = (uint32_t) (start_bit_ & 31);
    buf >>= skip;
    nbits -= skip;
    Refill();
  }
  JXLB_HD void Refill() {
    if (nbits <= 32) {
      uint32_t w = widx < wend ? words[widx] : 0u;
      buf |= (uint64_t) w << nbits;
      nbits += 32;
      ++widx;
    }
  }
  // All of Peek/Consume/Read require n <= 32.
  JXLB_HD uint32_t Peek(uint32_t n) const { return (uint32_t) (buf & ((1ull << n) - 1)); }
  JXLB_HD void Consume(uint32_t n) {
    buf >>= n;
    nbits -= n;
  }
  JXLB_HD uint32_t Read(uint32_t n) {
    Refill();
    uint32_t v = Peek(n);
    Consume(n);
    return v;
  }
  JXLB_HD uint32_t ReadBit() { return Read(1); }
  JXLB_HD uint64_t Position() const { return (uint64_t) widx * 32 - nbits; }  // absolute bit position
  JXLB_HD bool Overrun() const { return Position() > end_bit; }
  JXLB_HD void AlignToByte() {
    uint32_t r = (uint32_t) (Position() & 7);
    if (r) {
      Refill();
      Consume(8 - r);
    }
  }
  // Moves to an absolute bit position (used by the host parser between sections).
  JXLB_HD void Seek(uint64_t bitpos) {
    const uint8_t* base = reinterpret_cast<const uint8_t*>(words);
    uint64_t bytes = (uint64_t) wend << 2;
    uint64_t e = end_bit;
    Init(base, bytes, bitpos, e);
  }

  // ---- JPEG XL header field codings (App. B notation) -- host-side parsing, also usable on device.
  // U32 with four (base, extra_bits) options.
  JXLB_HD uint32_t U32(uint32_t b0, uint32_t n0, uint32_t b1, uint32_t n1, uint32_t b2, uint32_t n2, uint32_t b3, uint32_t n3) {
    uint32_t sel = Read(2);
    uint32_t b = sel == 0 ? b0 : sel == 1 ? b1 : sel == 2 ? b2 : b3;
    uint32_t n = sel == 0 ? n0 : sel == 1 ? n1 : sel == 2 ? n2 : n3;
    return b + (n ? Read(n) : 0u);
  }
  JXLB_HD uint64_t U64() {
    uint32_t sel = Read(2);
    if (sel == 0) return 0;
    if (sel == 1) return 1 + Read(4);
    if (sel == 2) return 17 + Read(8);
    uint64_t v = Read(12);
    uint32_t shift = 12;
    while (Read(1)) {
      if (shift == 60) {
        v |= (uint64_t) Read(4) << shift;
        break;
      }
      v |= (uint64_t) Read(8) << shift;
      shift += 8;
    }
    return v;
  }
  JXLB_HD uint32_t Enum() { return U32(0, 0, 1, 0, 2, 4, 18, 6); }
  JXLB_HD float F16() {
    uint32_t h = Read(16);
    uint32_t sign = h >> 15, exp = (h >> 10) & 31, mant = h & 1023;
    float v;
    if (exp == 0) {
      v = (float) mant * (1.0f / 16777216.0f);  // subnormal: mant * 2^-24
    } else {
      // (1 + mant/1024) * 2^(exp-15); exp == 31 (inf/nan) is invalid in JPEG XL headers and treated as finite here
      float m = 1.0f + (float) mant * (1.0f / 1024.0f);
      int e = (int) exp - 15;
      float s = 1.0f;
      for (int i = 0; i < (e < 0 ? -e : e); ++i) s *= 2.0f;
      v = e < 0 ? m / s : m * s;
    }
    return sign ? -v : v;
  }
};

}  // namespace jxlb
