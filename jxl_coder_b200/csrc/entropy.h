// JPEG XL entropy decoder (ANS with alias tables / Brotli-style prefix codes, hybrid-uint, LZ77, context maps),
// written once for host and device (see hd.h).  It replaces, for this path, what libjxl 0.12.0 does behind the
// reference's DecodeJpegXlOneShot (/root/reference/jxlcoder/src/main/cpp/interop/JxlDecoding.cpp:74-175 drives
// JxlDecoderProcessInput; libjxl's source is not in the reference tree).  Format digest: SURVEY.md App. B.4.
//
// A parsed code lives in a caller-provided arena as a position-independent blob (CodeHeader + tables addressed by
// offsets), so the host can parse the frame-global codes once and copy them to HBM verbatim, while streams that
// carry *local* codes (LF groups of streaming-encoded files) parse theirs on the device into per-stream scratch.
#pragma once
#include "bitreader.h"

namespace jxlb {

struct HybridCfg {
  uint8_t split_exp, msb, lsb, pad;
};

// One alias-table bucket (8 bytes, fetched with a single 64-bit load).
struct AliasEntry {
  uint8_t cutoff;      // positions >= cutoff in the bucket map to `right`
  uint8_t right;       // symbol for the upper part of the bucket
  uint16_t freq0;      // frequency of symbol == bucket index
  uint16_t offset1;    // offset to add to pos for the upper part
  uint16_t freq1;      // frequency of `right`
};

// Canonical prefix code of one cluster: counts per length and symbols sorted by (length, value).
struct PrefixCluster {
  uint16_t count[16];   // count[l] = number of codes of length l (1..15); count[0] = 1 marks a single-symbol code
  uint32_t syms_off;    // arena offset of uint16_t sorted symbols
  uint32_t single_sym;  // the symbol when count[0] == 1
};

struct CodeHeader {
  uint32_t num_ctx;       // contexts incl. the LZ77 distance context
  uint32_t num_clusters;
  uint8_t lz77, use_prefix, log_alpha, pad;
  uint32_t min_symbol, min_length;
  HybridCfg lz_len_cfg;
  uint32_t ctx_map_off;   // uint8_t[num_ctx]
  uint32_t cfg_off;       // HybridCfg[num_clusters]
  uint32_t table_off;     // AliasEntry[num_clusters << log_alpha]  or  PrefixCluster[num_clusters]
  uint32_t total_bytes;   // blob size (header included), multiple of 16
};

// Bump allocator over a byte buffer; offsets are relative to `base`.
struct Arena {
  uint8_t* base;
  uint32_t size;
  uint32_t used;
  JXLB_HD void Init(uint8_t* b, uint32_t s) {
    base = b;
    size = s;
    used = 0;
  }
  // returns offset or 0xFFFFFFFF
  JXLB_HD uint32_t Alloc(uint32_t bytes, uint32_t align = 16) {
    uint32_t o = (used + align - 1) & ~(align - 1);
    if (o > size || bytes > size - o) return 0xFFFFFFFFu;
    used = o + bytes;
    return o;
  }
};

static constexpr uint32_t kAnsTabBits = 12;
static constexpr uint32_t kAnsTabSize = 1u << kAnsTabBits;
static constexpr uint32_t kAnsSignature = 0x13;
static constexpr uint32_t kLz77WindowLog = 20;

JXLB_HD int ReadHybridCfg(BitReader& br, uint32_t log_alpha, HybridCfg* c) {
  c->split_exp = (uint8_t) br.Read(CeilLog2(log_alpha + 1));
  c->msb = c->lsb = c->pad = 0;
  if (c->split_exp > log_alpha) return kErrBadStream;
  if (c->split_exp != log_alpha) {
    c->msb = (uint8_t) br.Read(CeilLog2(c->split_exp + 1u));
    if (c->msb > c->split_exp) return kErrBadStream;
    c->lsb = (uint8_t) br.Read(CeilLog2(c->split_exp - c->msb + 1u));
    if ((uint32_t) c->msb + c->lsb > c->split_exp) return kErrBadStream;
  }
  return kOk;
}

JXLB_HD uint32_t ReadVarLen8(BitReader& br) {
  if (br.Read(1)) {
    uint32_t n = br.Read(3);
    return n == 0 ? 1u : br.Read(n) + (1u << n);
  }
  return 0;
}
JXLB_HD uint32_t ReadVarLen16(BitReader& br) {
  if (br.Read(1)) {
    uint32_t n = br.Read(4);
    return n == 0 ? 1u : br.Read(n) + (1u << n);
  }
  return 0;
}

// Fixed prefix code for the log-counts of an ANS histogram, on a 7-bit peek: low nibble pairs (nbits, symbol).
// (SURVEY.md App. B.4 step 5 and B.8 pitfall 1: entries 64.. repeat 0..63 except index 65 -> (7, 13).)
JXLB_HD void LogCountLookup(uint32_t idx7, uint32_t* nbits, uint32_t* sym) {
  // Decoded arithmetically instead of a 128-entry table (keeps it out of device local memory).
  uint32_t low3 = idx7 & 7;
  switch (low3) {
    case 0: *nbits = 3; *sym = 10; return;
    case 2: *nbits = 3; *sym = 7; return;
    case 4: *nbits = 3; *sym = 6; return;
    case 5: *nbits = 3; *sym = 8; return;
    case 6: *nbits = 3; *sym = 9; return;
    default: break;
  }
  uint32_t low4 = idx7 & 15;
  switch (low4) {
    case 3: *nbits = 4; *sym = 3; return;
    case 7: *nbits = 4; *sym = 5; return;
    case 9: *nbits = 4; *sym = 4; return;
    case 11: *nbits = 4; *sym = 1; return;
    case 15: *nbits = 4; *sym = 2; return;
    default: break;
  }
  // low4 == 1: 5+ bits
  uint32_t low5 = idx7 & 31;
  if (low5 == 17) { *nbits = 5; *sym = 0; return; }
  uint32_t low6 = idx7 & 63;  // low5 == 1
  if (low6 == 33) { *nbits = 6; *sym = 11; return; }
  // low6 == 1
  *nbits = 7;
  *sym = (idx7 & 64) ? 13 : 12;
}

// Reads one ANS histogram (counts summing to 4096) into counts[0..*alphabet).  counts must hold 256+8 entries.
JXLB_HD int ReadAnsHistogram(BitReader& br, uint32_t log_alpha, uint16_t* counts, uint32_t* alphabet) {
  const uint32_t table_size = 1u << log_alpha;
  if (br.Read(1)) {  // simple: one or two symbols
    uint32_t ns = br.Read(1) + 1;
    uint32_t s0 = ReadVarLen8(br), s1 = 0;
    if (ns == 2) s1 = ReadVarLen8(br);
    uint32_t mx = (ns == 2 && s1 > s0) ? s1 : s0;
    if (mx >= table_size) return kErrBadStream;
    for (uint32_t i = 0; i <= mx; ++i) counts[i] = 0;
    if (ns == 1) {
      counts[s0] = kAnsTabSize;
    } else {
      if (s0 == s1) return kErrBadStream;
      counts[s0] = (uint16_t) br.Read(12);
      counts[s1] = (uint16_t) (kAnsTabSize - counts[s0]);
    }
    *alphabet = mx + 1;
    return kOk;
  }
  if (br.Read(1)) {  // flat
    uint32_t n = ReadVarLen8(br) + 1;
    if (n > table_size) return kErrBadStream;
    for (uint32_t i = 0; i < n; ++i) counts[i] = (uint16_t) (kAnsTabSize / n + (i < kAnsTabSize % n ? 1 : 0));
    *alphabet = n;
    return kOk;
  }
  uint32_t log = 0;
  while (log < 3 && br.Read(1)) ++log;
  uint32_t shift = (br.Read(log) | (1u << log)) - 1;
  if (shift > kAnsTabBits + 1) return kErrBadStream;
  uint32_t length = ReadVarLen8(br) + 3;
  if (length > table_size) return kErrBadStream;
  uint8_t logcounts[264];
  uint16_t same[264];
  int omit_log = -1, omit_pos = -1;
  for (uint32_t i = 0; i < length; ++i) {
    same[i] = 0;
    logcounts[i] = 0;
  }
  for (uint32_t i = 0; i < length; ++i) {
    br.Refill();
    uint32_t nb, sym;
    LogCountLookup(br.Peek(7), &nb, &sym);
    br.Consume(nb);
    logcounts[i] = (uint8_t) sym;
    if (sym == 13) {  // RLE
      uint32_t rl = ReadVarLen8(br);
      same[i] = (uint16_t) (rl + 5);
      i += rl + 3;
      continue;
    }
    if ((int) sym > omit_log) {
      omit_log = (int) sym;
      omit_pos = (int) i;
    }
  }
  if (omit_pos < 0) return kErrBadStream;
  if ((uint32_t) omit_pos + 1 < length && logcounts[omit_pos + 1] == 13) return kErrBadStream;
  uint32_t total = 0, prev = 0, numsame = 0;
  for (uint32_t i = 0; i < length; ++i) {
    uint32_t c = 0;
    if (same[i]) {
      numsame = same[i] - 1u;
      prev = i > 0 ? counts[i - 1] : 0;
    }
    if (numsame > 0) {
      c = prev;
      --numsame;
    } else {
      uint32_t code = logcounts[i];
      if ((int) i == omit_pos || code == 0) {
        counts[i] = 0;
        continue;
      } else if (code == 1) {
        c = 1;
      } else {
        uint32_t lc = code - 1;
        int bc = (int) shift - (int) ((kAnsTabBits - lc) >> 1);
        if (bc < 0) bc = 0;
        if ((uint32_t) bc > lc) bc = (int) lc;
        c = (1u << lc) + (br.Read((uint32_t) bc) << (lc - (uint32_t) bc));
      }
    }
    counts[i] = (uint16_t) c;
    total += c;
    if (total > kAnsTabSize) return kErrBadStream;
  }
  if (total >= kAnsTabSize) return kErrBadStream;
  counts[omit_pos] = (uint16_t) (kAnsTabSize - total);
  *alphabet = length;
  return kOk;
}

// Builds the alias table for one histogram (App. B.4 step 5).
JXLB_HD void BuildAliasTable(const uint16_t* counts, uint32_t alphabet, uint32_t log_alpha, AliasEntry* table) {
  const uint32_t T = 1u << log_alpha;
  const uint32_t es = kAnsTabSize >> log_alpha;  // bucket size
  while (alphabet > 0 && counts[alphabet - 1] == 0) --alphabet;
  for (uint32_t s = 0; s < alphabet; ++s) {
    if (counts[s] == kAnsTabSize) {
      for (uint32_t i = 0; i < T; ++i) {
        table[i].cutoff = 0;
        table[i].right = (uint8_t) s;
        table[i].freq0 = 0;
        table[i].offset1 = (uint16_t) (es * i);
        table[i].freq1 = (uint16_t) kAnsTabSize;
      }
      return;
    }
  }
  if (alphabet == 0) {  // empty histogram: behaves like a single symbol 0
    for (uint32_t i = 0; i < T; ++i) {
      table[i].cutoff = 0;
      table[i].right = 0;
      table[i].freq0 = 0;
      table[i].offset1 = (uint16_t) (es * i);
      table[i].freq1 = (uint16_t) kAnsTabSize;
    }
    return;
  }
  uint16_t cut[256];
  uint16_t off1[256];
  uint8_t right[256];
  uint8_t under[256];
  uint8_t over[256];
  uint32_t nu = 0, no = 0;
  for (uint32_t i = 0; i < T; ++i) {
    cut[i] = i < alphabet ? counts[i] : 0;
    off1[i] = 0;
    right[i] = 0;
  }
  for (uint32_t i = 0; i < alphabet; ++i) {
    if (cut[i] > es) over[no++] = (uint8_t) i;
    else if (cut[i] < es) under[nu++] = (uint8_t) i;
  }
  for (uint32_t i = alphabet; i < T; ++i) under[nu++] = (uint8_t) i;
  while (no > 0) {
    uint32_t o = over[--no];
    uint32_t u = under[--nu];
    uint32_t by = es - cut[u];
    cut[o] = (uint16_t) (cut[o] - by);
    right[u] = (uint8_t) o;
    off1[u] = cut[o];
    if (cut[o] < es) under[nu++] = (uint8_t) o;
    else if (cut[o] > es) over[no++] = (uint8_t) o;
  }
  for (uint32_t i = 0; i < T; ++i) {
    AliasEntry e;
    if (cut[i] == es) {
      e.right = (uint8_t) i;
      e.offset1 = 0;
      e.cutoff = 0;
    } else {
      e.right = right[i];
      e.offset1 = (uint16_t) (off1[i] - cut[i]);
      e.cutoff = (uint8_t) cut[i];
    }
    e.freq0 = i < alphabet ? counts[i] : 0;
    e.freq1 = e.right < alphabet ? counts[e.right] : 0;
    table[i] = e;
  }
}

// ---- prefix codes (RFC 7932 §3.4 / §3.5) ---------------------------------------------------------------------------
// Turns code lengths (0..15) into the canonical decode structure.  lengths is a scratch array of `alphabet` bytes.
JXLB_HD int BuildPrefixCluster(const uint8_t* lengths, uint32_t alphabet, PrefixCluster* pc, uint16_t* syms) {
  for (int l = 0; l < 16; ++l) pc->count[l] = 0;
  uint32_t n = 0;
  for (uint32_t s = 0; s < alphabet; ++s)
    if (lengths[s]) {
      pc->count[lengths[s]]++;
      ++n;
    }
  uint32_t k = 0;
  for (uint32_t l = 1; l < 16; ++l)
    for (uint32_t s = 0; s < alphabet; ++s)
      if (lengths[s] == l) syms[k++] = (uint16_t) s;
  (void) n;
  return kOk;
}

JXLB_HD uint32_t ReadPrefixSymbol(BitReader& br, const PrefixCluster* pc, const uint16_t* syms) {
  if (pc->count[0]) return pc->single_sym;
  br.Refill();
  uint32_t code = 0, first = 0, index = 0;
  uint32_t bits = br.Peek(15);
  for (uint32_t len = 1; len < 16; ++len) {
    code |= bits & 1;
    bits >>= 1;
    uint32_t count = pc->count[len];
    if (code < first + count) {
      br.Consume(len);
      return syms[index + (code - first)];
    }
    index += count;
    first += count;
    first <<= 1;
    code <<= 1;
  }
  br.Consume(15);
  return 0;  // invalid code; caller's final checks catch corrupt streams
}

// Reads one prefix code definition for `alphabet` symbols.  `lengths` = scratch of alphabet bytes.
JXLB_HD int ReadPrefixCode(BitReader& br, uint32_t alphabet, uint8_t* lengths, PrefixCluster* pc, uint16_t* syms) {
  for (int l = 0; l < 16; ++l) pc->count[l] = 0;
  pc->single_sym = 0;
  if (alphabet == 1) {
    pc->count[0] = 1;
    return kOk;
  }
  for (uint32_t i = 0; i < alphabet; ++i) lengths[i] = 0;
  uint32_t hskip = br.Read(2);
  if (hskip == 1) {  // simple code: 1..4 symbols
    uint32_t mb = (uint32_t) FloorLog2(alphabet - 1) + 1;  // bits needed for alphabet-1
    uint32_t ns = br.Read(2) + 1;
    uint32_t s[4] = {0, 0, 0, 0};
    for (uint32_t i = 0; i < ns; ++i) {
      s[i] = br.Read(mb);
      if (s[i] >= alphabet) return kErrBadStream;
    }
    if (ns == 1) {
      pc->count[0] = 1;
      pc->single_sym = s[0];
      return kOk;
    }
    if (ns == 2) {
      lengths[s[0]] = 1;
      lengths[s[1]] = 1;
    } else if (ns == 3) {
      lengths[s[0]] = 1;
      lengths[s[1]] = 2;
      lengths[s[2]] = 2;
    } else {
      if (br.Read(1)) {
        lengths[s[0]] = 1;
        lengths[s[1]] = 2;
        lengths[s[2]] = 3;
        lengths[s[3]] = 3;
      } else {
        for (int i = 0; i < 4; ++i) lengths[s[i]] = 2;
      }
    }
    return BuildPrefixCluster(lengths, alphabet, pc, syms);
  }
  // complex code: code-length code first
  const uint8_t kOrder[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};
  uint8_t cl[18];
  for (int i = 0; i < 18; ++i) cl[i] = 0;
  int space = 32;
  uint32_t num = 0;
  for (uint32_t i = hskip; i < 18; ++i) {
    br.Refill();
    uint32_t p = br.Peek(4);
    uint32_t nb, v;
    // fixed code: 00->0, 10(lsb first: bits '01')... decoded as in RFC 7932 §3.5 on the 4 peeked bits
    if ((p & 3) == 0) { nb = 2; v = 0; }
    else if ((p & 3) == 1) { nb = 2; v = 4; }
    else if ((p & 3) == 2) { nb = 2; v = 3; }
    else if ((p & 7) == 3) { nb = 3; v = 2; }
    else if (p == 7) { nb = 4; v = 1; }
    else { nb = 4; v = 5; }
    br.Consume(nb);
    cl[kOrder[i]] = (uint8_t) v;
    if (v) {
      space -= 32 >> v;
      ++num;
    }
    if (space <= 0) break;
  }
  if (!(num == 1 || space == 0)) return kErrBadStream;
  // canonical code over the 18 code-length symbols
  PrefixCluster clpc;
  uint16_t clsyms[18];
  BuildPrefixCluster(cl, 18, &clpc, clsyms);
  clpc.single_sym = 0;
  int single_cl = -1;
  if (num == 1)
    for (int k = 0; k < 18; ++k)
      if (cl[k]) single_cl = k;
  uint32_t i = 0, prev = 8, rep = 0, rep_len = 0;
  int sp = 32768;
  while (i < alphabet && sp > 0) {
    uint32_t c = single_cl >= 0 ? (uint32_t) single_cl : ReadPrefixSymbol(br, &clpc, clsyms);
    if (c < 16) {
      rep = 0;
      lengths[i++] = (uint8_t) c;
      if (c) {
        prev = c;
        sp -= 32768 >> c;
      }
    } else {
      uint32_t extra = c - 14;
      uint32_t new_len = c == 16 ? prev : 0;
      if (rep_len != new_len) {
        rep = 0;
        rep_len = new_len;
      }
      uint32_t old = rep;
      if (rep > 0) rep = (rep - 2) << extra;
      rep += br.Read(extra) + 3;
      uint32_t delta = rep - old;
      if (i + delta > alphabet) return kErrBadStream;
      for (uint32_t k = 0; k < delta; ++k) lengths[i++] = (uint8_t) rep_len;
      if (rep_len) sp -= (int) (delta << (15 - rep_len));
    }
  }
  if (sp != 0) return kErrBadStream;
  return BuildPrefixCluster(lengths, alphabet, pc, syms);
}

// ---- code access -----------------------------------------------------------------------------------------------------
// Resolved view of a code blob, kept in registers while decoding a stream.
struct CodeView {
  const uint8_t* ctx_map;
  const HybridCfg* cfg;
  const AliasEntry* alias;
  const PrefixCluster* prefix;
  const uint8_t* blob;
  uint32_t num_ctx;
  uint32_t min_symbol, min_length;
  HybridCfg lz_len_cfg;
  uint8_t lz77, use_prefix, log_alpha, log_entry;  // log_entry = 12 - log_alpha
  JXLB_HD void Bind(const uint8_t* b) {
    const CodeHeader* h = reinterpret_cast<const CodeHeader*>(b);
    blob = b;
    ctx_map = b + h->ctx_map_off;
    cfg = reinterpret_cast<const HybridCfg*>(b + h->cfg_off);
    alias = reinterpret_cast<const AliasEntry*>(b + h->table_off);
    prefix = reinterpret_cast<const PrefixCluster*>(b + h->table_off);
    num_ctx = h->num_ctx;
    min_symbol = h->min_symbol;
    min_length = h->min_length;
    lz_len_cfg = h->lz_len_cfg;
    lz77 = h->lz77;
    use_prefix = h->use_prefix;
    log_alpha = h->log_alpha;
    log_entry = (uint8_t) (h->use_prefix ? 0 : kAnsTabBits - h->log_alpha);
  }
};

// Per-stream decoder state.
struct SymbolReader {
  uint32_t state;          // ANS state
  uint32_t num_to_copy, copy_pos, num_decoded;  // LZ77
  uint32_t* window;        // LZ77 window (1 << kLz77WindowLog entries, or smaller power of two: window_mask + 1)
  uint32_t window_mask;
  JXLB_HD void Begin(const CodeView& c, BitReader& br, uint32_t* win, uint32_t win_mask) {
    state = c.use_prefix ? (kAnsSignature << 16) : br.Read(32);
    num_to_copy = copy_pos = num_decoded = 0;
    window = win;
    window_mask = win_mask;
  }
  JXLB_HD bool FinalStateOk() const { return state == (kAnsSignature << 16); }
};

JXLB_HD uint32_t ReadToken(const CodeView& c, SymbolReader& r, BitReader& br, uint32_t cluster) {
  if (c.use_prefix) {
    const PrefixCluster* pc = c.prefix + cluster;
    return ReadPrefixSymbol(br, pc, reinterpret_cast<const uint16_t*>(c.blob + pc->syms_off));
  }
  uint32_t res = r.state & (kAnsTabSize - 1);
  uint32_t i = res >> c.log_entry;
  uint32_t pos = res & ((1u << c.log_entry) - 1);
  AliasEntry e = c.alias[(cluster << c.log_alpha) + i];
  bool hi = pos >= e.cutoff;
  uint32_t sym = hi ? e.right : i;
  uint32_t off = hi ? e.offset1 + pos : pos;
  uint32_t freq = hi ? e.freq1 : e.freq0;
  r.state = freq * (r.state >> kAnsTabBits) + off;
  if (r.state < (1u << 16)) {
    br.Refill();
    r.state = (r.state << 16) | br.Peek(16);
    br.Consume(16);
  }
  return sym;
}

JXLB_HD uint32_t DecodeHybridValue(const HybridCfg cfg, uint32_t token, BitReader& br) {
  uint32_t split = 1u << cfg.split_exp;
  if (token < split) return token;
  uint32_t in_token = (uint32_t) cfg.msb + cfg.lsb;
  uint32_t nbits = cfg.split_exp - in_token + ((token - split) >> in_token);
  if (nbits > 31) nbits = 31;  // corrupt stream guard; valid streams stay below
  uint32_t low = token & ((1u << cfg.lsb) - 1);
  token >>= cfg.lsb;
  uint32_t bits = br.Read(nbits);
  return ((((1u << cfg.msb) | (token & ((1u << cfg.msb) - 1))) << nbits | bits) << cfg.lsb) | low;
}

// 120 special (dx, dy) LZ77 distances for streams with a 2-D distance multiplier (JPEG XL spec table).
JXLB_HD int SpecialDistance(uint32_t i, int mult) {
  const int8_t k[120][2] = {
      {0, 1}, {1, 0}, {1, 1}, {-1, 1}, {0, 2}, {2, 0}, {1, 2}, {-1, 2}, {2, 1}, {-2, 1}, {2, 2}, {-2, 2}, {0, 3}, {3, 0}, {1, 3},
      {-1, 3}, {3, 1}, {-3, 1}, {2, 3}, {-2, 3}, {3, 2}, {-3, 2}, {0, 4}, {4, 0}, {1, 4}, {-1, 4}, {4, 1}, {-4, 1}, {3, 3}, {-3, 3},
      {2, 4}, {-2, 4}, {4, 2}, {-4, 2}, {0, 5}, {3, 4}, {-3, 4}, {4, 3}, {-4, 3}, {5, 0}, {1, 5}, {-1, 5}, {5, 1}, {-5, 1}, {2, 5},
      {-2, 5}, {5, 2}, {-5, 2}, {4, 4}, {-4, 4}, {3, 5}, {-3, 5}, {5, 3}, {-5, 3}, {0, 6}, {6, 0}, {1, 6}, {-1, 6}, {6, 1}, {-6, 1},
      {2, 6}, {-2, 6}, {6, 2}, {-6, 2}, {4, 5}, {-4, 5}, {5, 4}, {-5, 4}, {3, 6}, {-3, 6}, {6, 3}, {-6, 3}, {0, 7}, {7, 0}, {1, 7},
      {-1, 7}, {5, 5}, {-5, 5}, {7, 1}, {-7, 1}, {4, 6}, {-4, 6}, {6, 4}, {-6, 4}, {2, 7}, {-2, 7}, {7, 2}, {-7, 2}, {3, 7}, {-3, 7},
      {7, 3}, {-7, 3}, {5, 6}, {-5, 6}, {6, 5}, {-6, 5}, {8, 0}, {4, 7}, {-4, 7}, {7, 4}, {-7, 4}, {8, 1}, {8, 2}, {6, 6}, {-6, 6},
      {8, 3}, {5, 7}, {-5, 7}, {7, 5}, {-7, 5}, {8, 4}, {6, 7}, {-6, 7}, {7, 6}, {-7, 6}, {8, 5}, {7, 7}, {-7, 7}, {8, 6}, {8, 7}};
  int d = k[i][0] + mult * k[i][1];
  return d < 1 ? 1 : d;
}

// Reads one integer for context `ctx` (App. B.4 steps 6-8).  dist_mult = 0 for non-image streams.
JXLB_HD uint32_t ReadHybridUint(const CodeView& c, SymbolReader& r, BitReader& br, uint32_t ctx, int dist_mult = 0) {
  if (c.lz77) {
    if (r.num_to_copy > 0) {
      uint32_t v = r.window[r.copy_pos++ & r.window_mask];
      --r.num_to_copy;
      r.window[r.num_decoded++ & r.window_mask] = v;
      return v;
    }
    uint32_t cluster = c.ctx_map[ctx];
    uint32_t token = ReadToken(c, r, br, cluster);
    if (token >= c.min_symbol) {
      uint32_t num = DecodeHybridValue(c.lz_len_cfg, token - c.min_symbol, br) + c.min_length;
      uint32_t dcluster = c.ctx_map[c.num_ctx - 1];
      uint32_t dtok = ReadToken(c, r, br, dcluster);
      uint32_t dist = DecodeHybridValue(c.cfg[dcluster], dtok, br);
      if (dist_mult == 0) dist += 1;
      else if (dist >= 120) dist -= 119;
      else dist = (uint32_t) SpecialDistance(dist, dist_mult);
      if (dist > r.num_decoded) dist = r.num_decoded;
      if (dist > (1u << kLz77WindowLog)) dist = 1u << kLz77WindowLog;
      r.copy_pos = r.num_decoded - dist;
      if (dist == 0) {
        // nothing decoded yet: libjxl copies zeros; emulate by making the window read zeros
        for (uint32_t k = 0; k < num && k <= r.window_mask; ++k) r.window[k & r.window_mask] = 0;
      }
      r.num_to_copy = num;
      // first copied value
      uint32_t v = r.window[r.copy_pos++ & r.window_mask];
      --r.num_to_copy;
      r.window[r.num_decoded++ & r.window_mask] = v;
      return v;
    }
    uint32_t v = DecodeHybridValue(c.cfg[cluster], token, br);
    r.window[r.num_decoded++ & r.window_mask] = v;
    return v;
  }
  uint32_t cluster = c.ctx_map[ctx];
  uint32_t token = ReadToken(c, r, br, cluster);
  return DecodeHybridValue(c.cfg[cluster], token, br);
}

// ---- code parsing ----------------------------------------------------------------------------------------------------
// kDepth: 0 = a stream's code, 1 = the code of a context map, 2 = the code of the 2-entry context map that an
// LZ77-enabled context-map code needs (the format's recursion ends there: n <= 2 forbids LZ77).
template <int kDepth>
JXLB_HD_NOINLINE int ParseCodeT(BitReader& br, uint32_t num_ctx, bool allow_lz77, Arena& arena, uint32_t* blob_off);

// Context map of `n` entries (App. B.4 step 2) written to out[0..n); *num_clusters = max + 1.
template <int kDepth>
JXLB_HD_NOINLINE int ReadContextMapT(BitReader& br, uint32_t n, uint8_t* out, uint32_t* num_clusters, Arena& arena) {
  uint32_t mx = 0;
  if (br.Read(1)) {  // simple
    uint32_t b = br.Read(2);
    for (uint32_t i = 0; i < n; ++i) {
      out[i] = (uint8_t) (b ? br.Read(b) : 0);
      if (out[i] > mx) mx = out[i];
    }
    *num_clusters = mx + 1;
    return kOk;
  }
  uint32_t use_mtf = br.Read(1);
  // nested single-context code parsed into temporary arena space (released afterwards)
  uint32_t saved = arena.used;
  uint32_t off;
  int st = ParseCodeT<kDepth + 1>(br, 1, n > 2, arena, &off);
  if (st != kOk) return st;
  CodeView cv;
  cv.Bind(arena.base + off);
  uint32_t* win = nullptr;
  uint32_t wmask = 0;
  if (cv.lz77) {
    // window sized to the number of symbols (power of two >= n)
    uint32_t wl = 1;
    while (wl < n) wl <<= 1;
    uint32_t wo = arena.Alloc(wl * 4, 16);
    if (wo == 0xFFFFFFFFu) return kErrScratch;
    win = reinterpret_cast<uint32_t*>(arena.base + wo);
    wmask = wl - 1;
  }
  SymbolReader sr;
  sr.Begin(cv, br, win, wmask);
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t v = ReadHybridUint(cv, sr, br, 0);
    if (v > 255) return kErrBadStream;
    out[i] = (uint8_t) v;
  }
  if (!sr.FinalStateOk()) return kErrBadStream;
  arena.used = saved;
  if (use_mtf) {
    uint8_t mtf[256];
    for (int i = 0; i < 256; ++i) mtf[i] = (uint8_t) i;
    for (uint32_t i = 0; i < n; ++i) {
      uint32_t idx = out[i];
      uint8_t v = mtf[idx];
      out[i] = v;
      for (uint32_t k = idx; k > 0; --k) mtf[k] = mtf[k - 1];
      mtf[0] = v;
    }
  }
  for (uint32_t i = 0; i < n; ++i)
    if (out[i] > mx) mx = out[i];
  *num_clusters = mx + 1;
  return kOk;
}

// Parses a code over num_ctx contexts into the arena; *blob_off = offset of its CodeHeader.
template <int kDepth>
JXLB_HD_NOINLINE int ParseCodeT(BitReader& br, uint32_t num_ctx, bool allow_lz77, Arena& arena, uint32_t* blob_off) {
  uint32_t hoff = arena.Alloc(sizeof(CodeHeader), 16);
  if (hoff == 0xFFFFFFFFu) return kErrScratch;
  // NB: the arena may not move, so pointers into it stay valid.
  CodeHeader* h = reinterpret_cast<CodeHeader*>(arena.base + hoff);
  h->lz77 = (uint8_t) br.Read(1);
  h->min_symbol = h->min_length = 0;
  h->lz_len_cfg = HybridCfg{0, 0, 0, 0};
  h->pad = 0;
  if (h->lz77) {
    if (!allow_lz77) return kErrBadStream;
    h->min_symbol = br.U32(224, 0, 512, 0, 4096, 0, 8, 15);
    h->min_length = br.U32(3, 0, 4, 0, 5, 2, 9, 8);
    int st = ReadHybridCfg(br, 8, &h->lz_len_cfg);
    if (st != kOk) return st;
    num_ctx += 1;
  }
  h->num_ctx = num_ctx;
  uint32_t cmo = arena.Alloc(num_ctx, 16);
  if (cmo == 0xFFFFFFFFu) return kErrScratch;
  h->ctx_map_off = cmo - hoff;
  uint8_t* ctx_map = arena.base + cmo;
  uint32_t num_clusters = 1;
  if (num_ctx > 1) {
    if constexpr (kDepth < 2) {
      int st = ReadContextMapT<kDepth>(br, num_ctx, ctx_map, &num_clusters, arena);
      if (st != kOk) return st;
    } else {
      return kErrBadStream;  // cannot happen in a valid stream
    }
  } else {
    ctx_map[0] = 0;
  }
  h->num_clusters = num_clusters;
  h->use_prefix = (uint8_t) br.Read(1);
  h->log_alpha = (uint8_t) (h->use_prefix ? 15 : 5 + br.Read(2));
  uint32_t cfo = arena.Alloc(num_clusters * (uint32_t) sizeof(HybridCfg), 16);
  if (cfo == 0xFFFFFFFFu) return kErrScratch;
  h->cfg_off = cfo - hoff;
  HybridCfg* cfg = reinterpret_cast<HybridCfg*>(arena.base + cfo);
  for (uint32_t i = 0; i < num_clusters; ++i) {
    int st = ReadHybridCfg(br, h->log_alpha, &cfg[i]);
    if (st != kOk) return st;
  }
  if (h->use_prefix) {
    uint32_t to = arena.Alloc(num_clusters * (uint32_t) sizeof(PrefixCluster), 16);
    if (to == 0xFFFFFFFFu) return kErrScratch;
    h->table_off = to - hoff;
    PrefixCluster* pcs = reinterpret_cast<PrefixCluster*>(arena.base + to);
    // alphabet sizes first, then the codes
    uint32_t aso = arena.Alloc(num_clusters * 2, 16);
    if (aso == 0xFFFFFFFFu) return kErrScratch;
    uint16_t* asz = reinterpret_cast<uint16_t*>(arena.base + aso);
    for (uint32_t i = 0; i < num_clusters; ++i) {
      uint32_t a = ReadVarLen16(br) + 1;
      if (a > (1u << 15)) return kErrBadStream;
      asz[i] = (uint16_t) a;
    }
    for (uint32_t i = 0; i < num_clusters; ++i) {
      uint32_t a = asz[i];
      uint32_t so = arena.Alloc(a * 2, 16);
      if (so == 0xFFFFFFFFu) return kErrScratch;
      uint32_t saved = arena.used;
      uint32_t lo = arena.Alloc(a, 16);
      if (lo == 0xFFFFFFFFu) return kErrScratch;
      pcs[i].syms_off = so - hoff;
      int st = ReadPrefixCode(br, a, arena.base + lo, &pcs[i], reinterpret_cast<uint16_t*>(arena.base + so));
      if (st != kOk) return st;
      arena.used = saved;  // release the lengths scratch
    }
  } else {
    uint32_t T = 1u << h->log_alpha;
    uint32_t to = arena.Alloc(num_clusters * T * (uint32_t) sizeof(AliasEntry), 16);
    if (to == 0xFFFFFFFFu) return kErrScratch;
    h->table_off = to - hoff;
    AliasEntry* tab = reinterpret_cast<AliasEntry*>(arena.base + to);
    for (uint32_t i = 0; i < num_clusters; ++i) {
      uint16_t counts[264];
      uint32_t alphabet = 0;
      int st = ReadAnsHistogram(br, h->log_alpha, counts, &alphabet);
      if (st != kOk) return st;
      BuildAliasTable(counts, alphabet, h->log_alpha, tab + (size_t) i * T);
    }
  }
  arena.used = (arena.used + 15u) & ~15u;
  h->total_bytes = arena.used - hoff;
  *blob_off = hoff;
  return kOk;
}

// `kNested` = true is kept as an alias for callers that parse a stream-level code.
template <bool kUnused>
JXLB_HD int ParseCode(BitReader& br, uint32_t num_ctx, bool allow_lz77, Arena& arena, uint32_t* blob_off) {
  return ParseCodeT<0>(br, num_ctx, allow_lz77, arena, blob_off);
}
JXLB_HD int ReadContextMap(BitReader& br, uint32_t n, uint8_t* out, uint32_t* num_clusters, Arena& arena) {
  return ReadContextMapT<0>(br, n, out, num_clusters, arena);
}

// Lehmer-coded permutation (App. B.3 / B.7).  perm[0..size) receives the permutation; `skip` leading entries are
// the identity.  temp: scratch of `size` uint32_t.
JXLB_HD uint32_t PermCtx(uint32_t v) {
  if (v == 0) return 0;
  uint32_t c = (uint32_t) FloorLog2(v) + 1;
  return c > 7 ? 7 : c;
}
// Lehmer digits temp[0 .. end) (temp[i] < size - i; digits from `end` on are zero) -> perm[0 .. size).  temp is
// overwritten.  Returns kOk or kErrBadStream.
JXLB_HD_NOINLINE int ExpandLehmer(uint32_t size, uint32_t end, uint32_t* perm, uint32_t* temp) {
  // Lehmer code -> permutation: perm[i] = the temp[i]-th element still unused.  Done with a Fenwick tree of "unused"
  // counts held in perm[] (k-th unused element and its removal in O(log size) each), the results written over the
  // consumed digits in temp[]: O(size log size) whatever the digits are -- they are attacker-controlled, and a shifting
  // list costs the sum of the digits (~10^10 moves for a crafted set of 65536-entry coefficient orders).
  for (uint32_t j = 1; j <= size; ++j) perm[j - 1] = j & (0u - j);  // tree[j] = lowbit(j): all elements unused
  uint32_t top = 1;
  while ((top << 1) <= size) top <<= 1;
  for (uint32_t i = 0; i < end; ++i) {
    uint32_t k = temp[i], pos = 0;
    for (uint32_t pw = top; pw; pw >>= 1) {
      const uint32_t nx = pos + pw;
      if (nx <= size && perm[nx - 1] <= k) {
        pos = nx;
        k -= perm[nx - 1];
      }
    }
    if (pos >= size) return kErrBadStream;  // cannot happen: temp[i] < size - i was checked above
    temp[i] = pos;                          // 0-based element
    for (uint32_t j = pos + 1; j <= size; j += j & (0u - j)) --perm[j - 1];
  }
  // back from tree to per-element "unused" flags (inverse of the linear-time build), then the unused elements in order
  for (uint32_t j = size; j >= 1; --j) {
    const uint32_t parent = j + (j & (0u - j));
    if (parent <= size) perm[parent - 1] -= perm[j - 1];
  }
  {
    uint32_t p = end;
    for (uint32_t e = 0; e < size; ++e)
      if (perm[e]) temp[p++] = e;
  }
  for (uint32_t i = 0; i < size; ++i) perm[i] = temp[i];
  return kOk;
}

JXLB_HD_NOINLINE int ReadPermutation(const CodeView& c, SymbolReader& r, BitReader& br, uint32_t size, uint32_t skip,
                                     uint32_t* perm, uint32_t* temp) {
  uint32_t end = ReadHybridUint(c, r, br, PermCtx(size));
  if (end > size - skip) return kErrBadStream;
  end += skip;
  uint32_t last = 0;
  for (uint32_t i = 0; i < size; ++i) temp[i] = 0;
  for (uint32_t i = skip; i < end; ++i) {
    uint32_t v = ReadHybridUint(c, r, br, PermCtx(last));
    if (v >= size - i) return kErrBadStream;
    temp[i] = v;
    last = v;
  }
  return ExpandLehmer(size, end, perm, temp);
}

}  // namespace jxlb
