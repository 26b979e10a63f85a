// Latency-tuned inner loops of the modular decoder for the two channel shapes that make up an LfGroup section of a
// libjxl-encoded VarDCT frame (host + device; functionally identical to DecodeModularChannels in modular.h):
//
//   * DecodeWpChannelTight     "WP-only" subtrees (every decision on the weighted-predictor error, every leaf
//                              Weighted / offset 0 / multiplier 1): the LF coefficient channels.
//   * DecodeSimpleChannelTight small subtrees over the local-neighbourhood properties 2..14 with non-weighted
//                              predictors: the HF-metadata channels (CfL factors, BlockInfo, EPF sharpness).
//
// A modular sub-stream is ONE serial dependency chain (sample -> prediction errors -> context -> ANS state -> next
// sample), so on the GPU the section time is (samples) x (latency of that chain in one lane).  ncu on the first
// version showed ~300 issued instructions per LF sample at one issue every 4 cycles (pure dependent-issue latency,
// no memory stalls).  These loops are therefore written to minimise the instruction count on the chain and to give
// the scheduler one straight basic block per sample:
//   - all neighbourhood state (samples x8, true errors, the four sub-predictor error sums) slides through registers;
//     the row above is read from padded scratch rows (shared memory on the device), so the inner loop has no edge tests
//     except the peeled last column; the four sub-predictor error rows are interleaved (int4 per column: one 128-bit
//     load and one 128-bit store per sample);
//   - the tree walk is a 1 KB LUT from the clamped error property straight to the entropy-code CLUSTER (context map
//     pre-applied), built once per subtree;
//   - the bit reader keeps the next 32-bit word prefetched in a register, so a refill never waits on memory;
//   - rows are kept x8 (the weighted predictor's fixed-point domain), removing the shifts from the chain.
// Format digest: SURVEY.md App. B.6.  Replaces, for this path, libjxl 0.12.0's modular decoder behind the reference's
// DecodeJpegXlOneShot (/root/reference/jxlcoder/src/main/cpp/interop/JxlDecoding.cpp:74-175).
#pragma once
#include "modular_fast.h"

namespace jxlb {

// Bit reader with a one-word lookahead; bit-position compatible with BitReader (widx = index of `nextw`).
struct TightBits {
  const uint32_t* words;
  uint64_t buf;
  uint32_t nbits, widx, wend, nextw;
  JXLB_HD void From(const BitReader& br) {
    words = br.words;
    buf = br.buf;
    nbits = br.nbits;
    widx = br.widx;
    wend = br.wend;
    nextw = widx < wend ? words[widx] : 0u;
  }
  JXLB_HD void To(BitReader& br) const {
    br.buf = buf;
    br.nbits = nbits;
    br.widx = widx;
  }
  JXLB_HD void Refill() {
    if (nbits <= 32) {
      buf |= (uint64_t) nextw << nbits;
      nbits += 32;
      ++widx;
      nextw = widx < wend ? words[widx] : 0u;
    }
  }
  JXLB_HD uint32_t Read(uint32_t n) {  // n <= 32
    Refill();
    const uint32_t v = (uint32_t) (buf & ((1ull << n) - 1));
    buf >>= n;
    nbits -= n;
    return v;
  }
};

// The parts of a CodeView the inner loops need, held in registers.
struct TightCode {
  const AliasEntry* alias;
  const HybridCfg* cfg;
  uint32_t log_alpha, log_entry;
};
JXLB_HD TightCode MakeTightCode(const CodeView& code) {
  TightCode t;
  t.alias = code.alias;
  t.cfg = code.cfg;
  t.log_alpha = code.log_alpha;
  t.log_entry = code.log_entry;
  return t;
}
// One alias bucket as two 32-bit words: {cutoff | right << 8 | freq0 << 16, offset1 | freq1 << 16} (one 64-bit load).
struct AliasWords {
  uint32_t x, y;
};
JXLB_HD AliasWords LoadAlias(const AliasEntry* p) {
  AliasWords w;
#ifdef __CUDA_ARCH__
  const uint2 v = *reinterpret_cast<const uint2*>(p);  // tables are 16-byte aligned, entries 8 bytes
  w.x = v.x;
  w.y = v.y;
#else
  w.x = (uint32_t) p->cutoff | ((uint32_t) p->right << 8) | ((uint32_t) p->freq0 << 16);
  w.y = (uint32_t) p->offset1 | ((uint32_t) p->freq1 << 16);
#endif
  return w;
}

// One ANS symbol of `cluster` (alias-table codes only) followed by its hybrid-uint extra bits.
JXLB_HD uint32_t TightReadUint(TightBits& tb, uint32_t& state, const TightCode& code, uint32_t cluster) {
  const uint32_t res = state & (kAnsTabSize - 1);
  const uint32_t bucket = res >> code.log_entry;
  const uint32_t pos = res & ((1u << code.log_entry) - 1);
  const AliasWords e = LoadAlias(code.alias + (cluster << code.log_alpha) + bucket);
  const HybridCfg cfg = code.cfg[cluster];
  const bool hi = pos >= (e.x & 0xFFu);
  const uint32_t sym = hi ? ((e.x >> 8) & 0xFFu) : bucket;
  const uint32_t off = hi ? (e.y & 0xFFFFu) + pos : pos;
  const uint32_t freq = hi ? (e.y >> 16) : (e.x >> 16);
  uint32_t s = freq * (state >> kAnsTabBits) + off;
  if (s < (1u << 16)) {
    tb.Refill();
    s = (s << 16) | (uint32_t) (tb.buf & 0xFFFFu);
    tb.buf >>= 16;
    tb.nbits -= 16;
  }
  state = s;
  const uint32_t split = 1u << cfg.split_exp;
  if (sym < split) return sym;
  const uint32_t in_token = (uint32_t) cfg.msb + cfg.lsb;
  uint32_t nbits = cfg.split_exp - in_token + ((sym - split) >> in_token);
  if (nbits > 31) nbits = 31;
  const uint32_t low = sym & ((1u << cfg.lsb) - 1);
  const uint32_t tok = sym >> cfg.lsb;
  const uint32_t bits = tb.Read(nbits);
  return ((((1u << cfg.msb) | (tok & ((1u << cfg.msb) - 1))) << nbits | bits) << cfg.lsb) | low;
}

struct alignas(16) Int4 {
  int32_t v[4];
};

// Scratch (int32 units) of the tight decoders for channels up to `w` wide; the base must be 16-byte aligned.
struct TightScratch {
  static constexpr uint32_t kHeadInts = 64 + 256;  // divtab | 1 KB cluster LUT
  JXLB_HD static uint32_t Pad(uint32_t w) { return (w + 4u) & ~3u; }  // >= w + 1, multiple of 4
  JXLB_HD static uint32_t WpInts(uint32_t w) { return kHeadInts + 12u * Pad(w); }
  JXLB_HD static uint32_t SimpleInts(uint32_t w) { return kHeadInts + 3u * Pad(w); }
};

JXLB_HD bool TightCodeOk(const CodeView& code) { return !code.lz77 && !code.use_prefix; }

JXLB_HD bool TightWpEligible(const SubtreeInfo& info, const CodeView& code, const WPHeader& wph, uint32_t xs, uint32_t cap_ints) {
  return info.wp_only && TightCodeOk(code) && wph.p3[3] == 0 && xs >= 2 && TightScratch::WpInts(xs) <= cap_ints;
}

// Decodes one WP-only channel.  `last_root` caches which subtree the LUT in scratch was built for (0xFFFFFFFF = none).
// kHints: (device) scratch is shared memory, the code tables / output plane / bitstream are global memory.
template <bool kHints>
JXLB_HD void DecodeWpChannelTight(TightBits& tb, uint32_t& ans_state, const CodeView& code_view, const TreeNode* tree,
                                  const SubtreeInfo& info, const WPHeader& wph, const ModChannel& c, int32_t* scratch,
                                  uint32_t* last_root) {
  const TightCode code = MakeTightCode(code_view);
  if (kHints) {
    JXLB_ASSUME_SHARED(scratch);
    JXLB_ASSUME_GLOBAL(code.alias);
    JXLB_ASSUME_GLOBAL(code.cfg);
    JXLB_ASSUME_GLOBAL(c.data);
    JXLB_ASSUME_GLOBAL(tb.words);
  }
  const uint32_t xs = c.w, P = TightScratch::Pad(xs);
  uint32_t* divtab = reinterpret_cast<uint32_t*>(scratch);
  uint8_t* lut = reinterpret_cast<uint8_t*>(scratch + 64);
  int32_t* err_rows = scratch + TightScratch::kHeadInts;       // [2][P]
  int32_t* smp_rows = err_rows + 2 * P;                        // [2][P], samples x8
  Int4* pe_rows = reinterpret_cast<Int4*>(smp_rows + 2 * P);   // [2][P]
  if (*last_root != info.root) {
    for (uint32_t k = 0; k < 64; ++k) divtab[k] = (1u << 24) / (k + 1);
    for (int v = -512; v < 512; ++v) {
      TreeNode nd = tree[info.root];
      while (nd.property >= 0) nd = tree[v > nd.split_or_offset ? nd.left_or_ctx : nd.right_or_mul];
      lut[v + 512] = code_view.ctx_map[nd.left_or_ctx];
    }
    *last_root = info.root;
  }
  for (uint32_t k = 0; k < 12u * P; ++k) err_rows[k] = 0;
  const int32_t hw0 = wph.w[0], hw1 = wph.w[1], hw2 = wph.w[2], hw3 = wph.w[3];
  const int32_t p1C = wph.p1C, p2C = wph.p2C, p3a = wph.p3[0], p3b = wph.p3[1], p3c = wph.p3[2], p3e = wph.p3[4];
  uint32_t state = ans_state;

  // One sample.  kFirst: row 0 (no row above: NE = N, and N / NW follow W); kEdge: last column (NE position == N position).
#define JXLB_WP_SAMPLE(kFirst, kEdge)                                                                         \
  {                                                                                                           \
    const Int4 pe_ne = (kEdge) ? pe_n : pe_prev[x + 1];                                                       \
    const int32_t err_ne = err_prev[x + 1];                                                                   \
    const int32_t NE8 = (kFirst) ? N8 : smp_prev[x + 1];                                                      \
    /* error-weighted sub-predictor weights */                                                                \
    uint32_t w[4];                                                                                            \
    const int32_t hw[4] = {hw0, hw1, hw2, hw3};                                                               \
    for (int i = 0; i < 4; ++i) {                                                                             \
      const uint32_t e = (uint32_t) pe_n.v[i] + (uint32_t) pe_ne.v[i] + (uint32_t) pe_nw.v[i];                \
      int shift = FloorLog2(e + 1) - 5;                                                                       \
      if (shift < 0) shift = 0;                                                                               \
      w[i] = 4 + (((uint32_t) hw[i] * divtab[e >> shift]) >> shift);                                          \
    }                                                                                                         \
    /* property 15: the neighbour error of largest magnitude (first of W, N, NW, NE on ties) -> cluster */    \
    int32_t pm = err_n;                                                                                       \
    {                                                                                                         \
      int32_t am = pm < 0 ? -pm : pm;                                                                         \
      const int32_t a1 = err_nw < 0 ? -err_nw : err_nw;                                                       \
      if (a1 > am) { pm = err_nw; am = a1; }                                                                  \
      const int32_t a2 = err_ne < 0 ? -err_ne : err_ne;                                                       \
      if (a2 > am) { pm = err_ne; am = a2; }                                                                  \
      const int32_t aw = err_w < 0 ? -err_w : err_w;                                                          \
      if (!(am > aw)) pm = err_w;                                                                             \
    }                                                                                                         \
    int32_t pv = pm < -512 ? -512 : pm > 511 ? 511 : pm;                                                      \
    const uint32_t cluster = lut[pv + 512];                                                                   \
    /* the four sub-predictions (x8) */                                                                       \
    int32_t sub[4];                                                                                           \
    sub[0] = W8 + NE8 - N8;                                                                                   \
    sub[1] = N8 - (((err_n + err_w + err_ne) * p1C) >> 5);                                                    \
    sub[2] = W8 - (((err_n + err_w + err_nw) * p2C) >> 5);                                                    \
    sub[3] = N8 - ((err_nw * p3a + err_n * p3b + err_ne * p3c + (NW8 - W8) * p3e) >> 5);                      \
    uint32_t ws = w[0] + w[1] + w[2] + w[3];                                                                  \
    const int lw = FloorLog2(ws) - 4;                                                                         \
    w[0] >>= lw;                                                                                              \
    w[1] >>= lw;                                                                                              \
    w[2] >>= lw;                                                                                              \
    w[3] >>= lw;                                                                                              \
    ws = w[0] + w[1] + w[2] + w[3];                                                                           \
    const int32_t sum = (int32_t) (ws >> 1) - 1 + sub[0] * (int32_t) w[0] + sub[1] * (int32_t) w[1] +         \
                        sub[2] * (int32_t) w[2] + sub[3] * (int32_t) w[3];                                    \
    int32_t pred = (int32_t) (((int64_t) sum * (int64_t) divtab[ws - 1]) >> 24);                              \
    if (((err_n ^ err_w) | (err_n ^ err_nw)) <= 0) {                                                          \
      int32_t mx = W8 > NE8 ? W8 : NE8;                                                                       \
      if (N8 > mx) mx = N8;                                                                                   \
      int32_t mn = W8 < NE8 ? W8 : NE8;                                                                       \
      if (N8 < mn) mn = N8;                                                                                   \
      if (pred > mx) pred = mx;                                                                               \
      if (pred < mn) pred = mn;                                                                               \
    }                                                                                                         \
    const uint32_t u = TightReadUint(tb, state, code, cluster);                                               \
    const int32_t val = UnpackSigned(u) + ((pred + 3) >> 3);                                                  \
    const int32_t v8 = val * 8;                                                                               \
    out_row[x] = val;                                                                                         \
    smp_cur[x] = v8;                                                                                          \
    const int32_t e_cur = pred - v8;                                                                          \
    err_cur[x] = e_cur;                                                                                       \
    Int4 en;                                                                                                  \
    for (int i = 0; i < 4; ++i) {                                                                             \
      int32_t d = sub[i] - v8;                                                                                \
      if (d < 0) d = -d;                                                                                      \
      en.v[i] = (d + 3) >> 3;                                                                                 \
    }                                                                                                         \
    pe_cur[x] = en;                                                                                           \
    for (int i = 0; i < 4; ++i) {                                                                             \
      pe_nw.v[i] = pe_n.v[i];                                                                                 \
      pe_n.v[i] = pe_ne.v[i] + en.v[i];                                                                       \
    }                                                                                                         \
    err_nw = err_n;                                                                                           \
    err_n = err_ne;                                                                                           \
    err_w = e_cur;                                                                                            \
    W8 = v8;                                                                                                  \
    if (kFirst) {                                                                                             \
      NW8 = v8;                                                                                               \
      N8 = v8;                                                                                                \
    } else {                                                                                                  \
      NW8 = N8;                                                                                               \
      N8 = NE8;                                                                                               \
    }                                                                                                         \
  }

  for (uint32_t y = 0; y < c.h; ++y) {
    int32_t* out_row = c.data + (size_t) y * c.stride;
    const uint32_t cur_o = (y & 1) * P, prv_o = P - cur_o;
    int32_t* err_cur = err_rows + cur_o;
    const int32_t* err_prev = err_rows + prv_o;
    int32_t* smp_cur = smp_rows + cur_o;
    const int32_t* smp_prev = smp_rows + prv_o;
    Int4* pe_cur = pe_rows + cur_o;
    const Int4* pe_prev = pe_rows + prv_o;
    Int4 pe_n = pe_prev[0], pe_nw = pe_n;
    int32_t err_n = err_prev[0], err_nw = err_n, err_w = 0;
    int32_t W8 = y > 0 ? smp_prev[0] : 0, N8 = W8, NW8 = W8;
    uint32_t x = 0;
    if (y == 0) {
      for (; x + 1 < xs; ++x) JXLB_WP_SAMPLE(true, false)
      JXLB_WP_SAMPLE(true, true)
    } else {
      for (; x + 1 < xs; ++x) JXLB_WP_SAMPLE(false, false)
      JXLB_WP_SAMPLE(false, true)
    }
    // pads read by the next row at its last column
    err_cur[xs] = err_cur[xs - 1];
    smp_cur[xs] = smp_cur[xs - 1];
  }
#undef JXLB_WP_SAMPLE
  ans_state = state;
}

// ---- small subtrees over neighbourhood properties ---------------------------------------------------------------------
struct SimpleNode {        // 16 bytes
  int32_t split_or_offset; // inner: split; leaf: offset
  uint16_t left, right;    // inner: children (indices into the compact array)
  int8_t property;         // -1 = leaf
  uint8_t predictor;       // leaf
  uint8_t cluster;         // leaf: entropy-code cluster (context map applied)
  uint8_t pad;
  uint32_t mul;            // leaf: multiplier
};
static constexpr uint32_t kMaxSimpleNodes = 63;

// Copies the subtree reachable from info.root into `out` (breadth-first).  Returns the node count, or 0 when the
// subtree does not qualify (too large, weighted predictor, previous-channel or WP properties).
JXLB_HD_NOINLINE uint32_t BuildSimpleSubtree(const TreeNode* tree, const SubtreeInfo& info, const CodeView& code, SimpleNode* out) {
  if (info.uses_wp || info.needs_general || !TightCodeOk(code)) return 0;
  uint32_t src[kMaxSimpleNodes];
  uint32_t n = 1;
  src[0] = info.root;
  for (uint32_t i = 0; i < n; ++i) {
    const TreeNode nd = tree[src[i]];
    SimpleNode s;
    s.split_or_offset = nd.split_or_offset;
    s.pad = 0;
    if (nd.property < 0) {
      if (nd.predictor == 6) return 0;
      s.property = -1;
      s.predictor = nd.predictor;
      s.cluster = code.ctx_map[nd.left_or_ctx];
      s.mul = nd.right_or_mul;
      s.left = s.right = 0;
    } else {
      if (nd.property < 2 || nd.property > 14) return 0;  // static properties were pruned; 15+ need other state
      if (n + 2 > kMaxSimpleNodes) return 0;
      s.property = (int8_t) nd.property;
      s.predictor = 0;
      s.cluster = 0;
      s.mul = 0;
      s.left = (uint16_t) n;
      src[n++] = nd.left_or_ctx;
      s.right = (uint16_t) n;
      src[n++] = nd.right_or_mul;
    }
    out[i] = s;
  }
  return n;
}

// Decodes one channel with a compact subtree (see BuildSimpleSubtree).  scratch: TightScratch::SimpleInts(c.w) ints with
// the node array at scratch + 64.
template <bool kHints>
JXLB_HD void DecodeSimpleChannelTight(TightBits& tb, uint32_t& ans_state, const CodeView& code_view, const ModChannel& c,
                                      int32_t* scratch) {
  const TightCode code = MakeTightCode(code_view);
  const SimpleNode* nodes = reinterpret_cast<const SimpleNode*>(scratch + 64);  // see the note in DecodeNwChannelTight
  if (kHints) {
    JXLB_ASSUME_SHARED(scratch);
    JXLB_ASSUME_GLOBAL(code.alias);
    JXLB_ASSUME_GLOBAL(code.cfg);
    JXLB_ASSUME_GLOBAL(c.data);
    JXLB_ASSUME_GLOBAL(tb.words);
  }
  const uint32_t xs = c.w, P = TightScratch::Pad(xs);
  int32_t* rows = scratch + TightScratch::kHeadInts;  // [3][P]: rotating current / N / NN rows
  int32_t* rbuf[3] = {rows, rows + P, rows + 2 * P};
  uint32_t state = ans_state;
  const SimpleNode root = nodes[0];
  for (uint32_t y = 0; y < c.h; ++y) {
    int32_t* out_row = c.data + (size_t) y * c.stride;
    int32_t* cur = rbuf[0];
    const int32_t* rN = rbuf[1];
    const int32_t* rNN = rbuf[2];
    int32_t W = y > 0 ? rN[0] : 0, N = W, NW = W, WW = W, prev9 = 0;
    for (uint32_t x = 0; x < xs; ++x) {
      const int32_t NE = (x + 1 < xs && y > 0) ? rN[x + 1] : N;
      SimpleNode nd = root;
      if (nd.property >= 0) {
        const int32_t NN = y > 1 ? rNN[x] : N;
        const int32_t p9 = (int32_t) ((int64_t) W + N - NW);
        do {
          int32_t v;
          switch (nd.property) {
            case 2: v = (int32_t) y; break;
            case 3: v = (int32_t) x; break;
            case 4: v = N < 0 ? -N : N; break;
            case 5: v = W < 0 ? -W : W; break;
            case 6: v = N; break;
            case 7: v = W; break;
            case 8: v = W - prev9; break;
            case 9: v = p9; break;
            case 10: v = W - NW; break;
            case 11: v = NW - N; break;
            case 12: v = N - NE; break;
            case 13: v = N - NN; break;
            default: v = W - WW; break;
          }
          nd = nodes[v > nd.split_or_offset ? nd.left : nd.right];
        } while (nd.property >= 0);
        prev9 = p9;
      }
      int64_t pred;
      if (nd.predictor == 0) {
        pred = 0;
      } else if (nd.predictor == 1) {
        pred = W;
      } else if (nd.predictor == 5) {
        pred = ClampedGradient(W, N, NW);
      } else {
        const int32_t NN = y > 1 ? rNN[x] : N;
        const int32_t NEE = (x + 2 < xs && y > 0) ? rN[x + 2] : NE;
        pred = PredictNoWp(nd.predictor, W, N, NW, NE, NN, WW, NEE);
      }
      const uint32_t u = TightReadUint(tb, state, code, nd.cluster);
      const int32_t val = (int32_t) ((int64_t) UnpackSigned(u) * (int64_t) nd.mul + nd.split_or_offset + pred);
      cur[x] = val;
      out_row[x] = val;
      WW = x >= 1 ? W : val;
      W = val;
      if (y > 0) {
        NW = N;
        N = NE;
      } else {
        NW = val;
        N = val;
      }
    }
    int32_t* t = rbuf[2];
    rbuf[2] = rbuf[1];
    rbuf[1] = rbuf[0];
    rbuf[0] = t;
  }
  ans_state = state;
}

// ---- subtrees over {y, N, W} only: context by table lookup ----------------------------------------------------------------
// The HF-metadata channels of libjxl-encoded frames (BlockInfo, EPF sharpness) use small trees that test only the row
// (property 2), N (6) and W (7) and predict Zero / W / Gradient.  A decision "v > split" is unchanged when v is clamped
// to [min split, max split + 1], so the whole walk collapses into one lookup indexed by the clamped (W, N); entries pack
// cluster | predictor << 8 | leaf << 16.  Rebuilt per row only when the tree tests y.
struct NwPlan {
  int32_t lo_w, lo_n;       // clamp ranges: [lo, lo + n - 1]
  uint32_t n_w, n_n;
  uint32_t uses_y, plain;   // plain: every leaf has multiplier 1 and offset 0
};
static constexpr uint32_t kNwLutEntries = 1024;
JXLB_HD uint32_t NwInts(uint32_t w) { return TightScratch::kHeadInts + 2u * TightScratch::Pad(w) + kNwLutEntries; }

// Examines a compact subtree (BuildSimpleSubtree output); false when it does not qualify.
JXLB_HD_NOINLINE bool PlanNwSubtree(const SimpleNode* nodes, uint32_t n, NwPlan* plan) {
  int32_t lo_w = 0x7FFFFFFF, hi_w = (int32_t) 0x80000000, lo_n = 0x7FFFFFFF, hi_n = (int32_t) 0x80000000;
  bool any_w = false, any_n = false;
  plan->uses_y = 0;
  plan->plain = 1;
  for (uint32_t i = 0; i < n; ++i) {
    const SimpleNode nd = nodes[i];
    if (nd.property < 0) {
      if (nd.predictor != 0 && nd.predictor != 1 && nd.predictor != 5) return false;
      if (nd.mul != 1 || nd.split_or_offset != 0) plan->plain = 0;
    } else if (nd.property == 2) {
      plan->uses_y = 1;
    } else if (nd.property == 6) {
      any_n = true;
      if (nd.split_or_offset < lo_n) lo_n = nd.split_or_offset;
      if (nd.split_or_offset > hi_n) hi_n = nd.split_or_offset;
    } else if (nd.property == 7) {
      any_w = true;
      if (nd.split_or_offset < lo_w) lo_w = nd.split_or_offset;
      if (nd.split_or_offset > hi_w) hi_w = nd.split_or_offset;
    } else {
      return false;
    }
  }
  if (!any_w) lo_w = hi_w = 0;
  if (!any_n) lo_n = hi_n = 0;
  if (hi_w > 0x3FFFFFFF || hi_n > 0x3FFFFFFF || lo_w < -0x3FFFFFFF || lo_n < -0x3FFFFFFF) return false;
  const int64_t nw = (int64_t) hi_w - lo_w + 2, nn = (int64_t) hi_n - lo_n + 2;
  if (nw * nn > (int64_t) kNwLutEntries) return false;
  plan->lo_w = lo_w;
  plan->lo_n = lo_n;
  plan->n_w = (uint32_t) nw;
  plan->n_n = (uint32_t) nn;
  return true;
}

JXLB_HD void BuildNwLut(const SimpleNode* nodes, const NwPlan& plan, uint32_t y, uint32_t* lut) {
  for (uint32_t iw = 0; iw < plan.n_w; ++iw)
    for (uint32_t in = 0; in < plan.n_n; ++in) {
      const int32_t wv = plan.lo_w + (int32_t) iw, nv = plan.lo_n + (int32_t) in;
      uint32_t k = 0;
      SimpleNode nd = nodes[0];
      while (nd.property >= 0) {
        const int32_t v = nd.property == 2 ? (int32_t) y : nd.property == 6 ? nv : wv;
        k = v > nd.split_or_offset ? nd.left : nd.right;
        nd = nodes[k];
      }
      lut[iw * plan.n_n + in] = (uint32_t) nd.cluster | ((uint32_t) nd.predictor << 8) | (k << 16);
    }
}

// scratch: NwInts(c.w) ints; the node array sits at scratch + 64 (as left by BuildSimpleSubtree).
template <bool kHints>
JXLB_HD void DecodeNwChannelTight(TightBits& tb, uint32_t& ans_state, const CodeView& code_view, const NwPlan& plan,
                                  const ModChannel& c, int32_t* scratch) {
  const TightCode code = MakeTightCode(code_view);
  // NOTE: hints are only ever applied to pointers that are shared / global on EVERY path that can produce them: nvcc
  // infers the address space of the underlying value, not of the program point (a select of a shared and a global
  // pointer hinted in one branch miscompiles the other).  Hence `nodes` is derived here, not passed in.
  const SimpleNode* nodes = reinterpret_cast<const SimpleNode*>(scratch + 64);
  if (kHints) {
    JXLB_ASSUME_SHARED(scratch);
    JXLB_ASSUME_GLOBAL(code.alias);
    JXLB_ASSUME_GLOBAL(code.cfg);
    JXLB_ASSUME_GLOBAL(c.data);
    JXLB_ASSUME_GLOBAL(tb.words);
  }
  const uint32_t xs = c.w, P = TightScratch::Pad(xs);
  int32_t* rows = scratch + TightScratch::kHeadInts;  // [2][P]: current / N rows
  uint32_t* lut = reinterpret_cast<uint32_t*>(rows + 2 * P);
  uint32_t state = ans_state;
  const int32_t lo_w = plan.lo_w, hi_w = plan.lo_w + (int32_t) plan.n_w - 1;
  const int32_t lo_n = plan.lo_n, hi_n = plan.lo_n + (int32_t) plan.n_n - 1;
  const uint32_t n_n = plan.n_n;
  if (!plan.uses_y) BuildNwLut(nodes, plan, 0, lut);
  for (uint32_t y = 0; y < c.h; ++y) {
    if (plan.uses_y) BuildNwLut(nodes, plan, y, lut);
    int32_t* out_row = c.data + (size_t) y * c.stride;
    int32_t* cur = rows + (y & 1) * P;
    const int32_t* rN = rows + ((y & 1) ^ 1) * P;
    int32_t W = y > 0 ? rN[0] : 0, N = W, NW = W;
    for (uint32_t x = 0; x < xs; ++x) {
      const int32_t n_next = (x + 1 < xs && y > 0) ? rN[x + 1] : N;  // the N of the next column (NE of this one)
      const int32_t wc = W < lo_w ? lo_w : W > hi_w ? hi_w : W;
      const int32_t nc = N < lo_n ? lo_n : N > hi_n ? hi_n : N;
      const uint32_t e = lut[(uint32_t) (wc - lo_w) * n_n + (uint32_t) (nc - lo_n)];
      const uint32_t predictor = (e >> 8) & 0xFFu;
      const int32_t pred = predictor == 0 ? 0 : predictor == 1 ? W : (int32_t) ClampedGradient(W, N, NW);
      const uint32_t u = TightReadUint(tb, state, code, e & 0xFFu);
      int32_t val;
      if (plan.plain) {
        val = UnpackSigned(u) + pred;
      } else {
        const SimpleNode lf = nodes[e >> 16];
        val = (int32_t) ((int64_t) UnpackSigned(u) * (int64_t) lf.mul + lf.split_or_offset + pred);
      }
      cur[x] = val;
      out_row[x] = val;
      W = val;
      if (y > 0) {
        NW = N;
        N = n_next;
      } else {
        NW = val;
        N = val;
      }
    }
  }
  ans_state = state;
}

// Decodes all channels of one modular sub-stream.  Streams whose every channel fits one of the tight loops (alias-table
// code without LZ77; WP-only or small neighbourhood subtrees) take them; anything else goes through the general decoder
// DecodeModularChannelsFast with identical results.  scratch: ModFastScratch::Ints(max w) ints (any memory);
// fast_scratch: optional low-latency memory (shared memory on the device) of fast_ints ints, 16-byte aligned.
#ifdef __CUDA_ARCH__
#define JXLB_DEV_HINTS true
#else
#define JXLB_DEV_HINTS false
#endif

JXLB_HD_NOINLINE int DecodeModularChannelsTight(BitReader& br_io, const ModularContext& mc, const WPHeader& wph, const ModChannel* ch,
                                                uint32_t nch, uint32_t stream_id, int32_t* scratch, uint32_t* lz77_window,
                                                uint32_t lz77_mask, int32_t* fast_scratch, uint32_t fast_ints) {
  constexpr uint32_t kMaxCh = 8;
  bool tight = TightCodeOk(mc.code) && nch <= kMaxCh;
#ifdef __CUDA_ARCH__
  // On the device the tight loops exist only in their hinted form: shared-memory scratch, everything else in HBM.
  if (!fast_scratch || !__isShared(fast_scratch) || !__isGlobal(mc.code.blob) || !__isGlobal(br_io.words) || !__isGlobal(scratch)) tight = false;
#endif
  uint8_t kind[kMaxCh];  // 0 = empty, 1 = WP-only, 2 = simple subtree
  SubtreeInfo infos[kMaxCh];
  for (uint32_t ci = 0; ci < nch && tight; ++ci) {
    kind[ci] = 0;
    if (!ch[ci].w || !ch[ci].h) continue;
#ifdef __CUDA_ARCH__
    if (!__isGlobal(ch[ci].data)) {
      tight = false;
      break;
    }
#endif
    infos[ci] = AnalyseSubtree(mc.tree, mc.num_nodes, ci, stream_id);
    const bool wp_fast = fast_scratch && TightScratch::WpInts(ch[ci].w) <= fast_ints;
    int32_t* base = wp_fast ? fast_scratch : scratch;
    if (TightWpEligible(infos[ci], mc.code, wph, ch[ci].w, 0xFFFFFFFFu) && (wp_fast || !JXLB_DEV_HINTS)) {
      kind[ci] = 1;
    } else if (BuildSimpleSubtree(mc.tree, infos[ci], mc.code, reinterpret_cast<SimpleNode*>(base + 64)) != 0) {
      kind[ci] = 2;
    } else {
      tight = false;
    }
  }
  if (!tight) return DecodeModularChannelsFast(br_io, mc, wph, ch, nch, stream_id, scratch, lz77_window, lz77_mask, fast_scratch, fast_ints);
  BitReader br = br_io;
  uint32_t state = br.Read(32);  // SymbolReader::Begin for an ANS code
  TightBits tb;
  tb.From(br);
  uint32_t last_root_fast = 0xFFFFFFFFu, last_root_slow = 0xFFFFFFFFu;
  for (uint32_t ci = 0; ci < nch; ++ci) {
    if (kind[ci] == 0) continue;
    const ModChannel c = ch[ci];
    if (kind[ci] == 1) {
      const bool fast = fast_scratch && TightScratch::WpInts(c.w) <= fast_ints;
      if (fast) DecodeWpChannelTight<JXLB_DEV_HINTS>(tb, state, mc.code, mc.tree, infos[ci], wph, c, fast_scratch, &last_root_fast);
#ifndef __CUDA_ARCH__
      else DecodeWpChannelTight<false>(tb, state, mc.code, mc.tree, infos[ci], wph, c, scratch, &last_root_slow);
#endif
    } else {
      // the compact node array goes to the head of whichever scratch the channel will use
      const bool nw_fast = fast_scratch && NwInts(c.w) <= fast_ints;
      const bool simple_fast = fast_scratch && TightScratch::SimpleInts(c.w) <= fast_ints;
      SimpleNode* nodes = reinterpret_cast<SimpleNode*>((nw_fast ? fast_scratch : scratch) + 64);
      const uint32_t nn = BuildSimpleSubtree(mc.tree, infos[ci], mc.code, nodes);
      NwPlan plan;
      if (PlanNwSubtree(nodes, nn, &plan)) {
        (nw_fast ? last_root_fast : last_root_slow) = 0xFFFFFFFFu;  // the node array overwrote the WP LUT
        if (nw_fast) DecodeNwChannelTight<JXLB_DEV_HINTS>(tb, state, mc.code, plan, c, fast_scratch);
        else DecodeNwChannelTight<false>(tb, state, mc.code, plan, c, scratch);
      } else {
        if (simple_fast != nw_fast) {
          nodes = reinterpret_cast<SimpleNode*>((simple_fast ? fast_scratch : scratch) + 64);
          BuildSimpleSubtree(mc.tree, infos[ci], mc.code, nodes);
        }
        (simple_fast ? last_root_fast : last_root_slow) = 0xFFFFFFFFu;
        if (simple_fast) DecodeSimpleChannelTight<JXLB_DEV_HINTS>(tb, state, mc.code, c, fast_scratch);
        else DecodeSimpleChannelTight<false>(tb, state, mc.code, c, scratch);
      }
    }
  }
  tb.To(br);
  br_io = br;
  if (state != (kAnsSignature << 16)) return kErrBadStream;
  if (br.Overrun()) return kErrTruncated;
  return kOk;
}

}  // namespace jxlb
