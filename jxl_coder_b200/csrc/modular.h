// Modular sub-bitstream decoder (MA tree, 14 predictors incl. the weighted predictor, previous-channel properties),
// host + device (see hd.h).  One call decodes one sub-stream serially — that is inherent to the format (every sample
// is predicted from already-decoded neighbours and the ANS state is sequential); the parallelism is across streams
// (LF groups, pass groups, images), which the kernels map to one warp-lane each.
// Replaces, for this path, libjxl 0.12.0's modular decoder behind the reference's DecodeJpegXlOneShot
// (/root/reference/jxlcoder/src/main/cpp/interop/JxlDecoding.cpp:74-175).  Format digest: SURVEY.md App. B.6.
#pragma once
#include "entropy.h"

namespace jxlb {

struct TreeNode {            // 16 bytes
  int16_t property;          // -1 = leaf
  uint8_t predictor;         // leaf
  uint8_t pad;
  int32_t split_or_offset;   // inner: split value; leaf: offset
  uint32_t left_or_ctx;      // inner: child taken when property > split; leaf: context id
  uint32_t right_or_mul;     // inner: other child; leaf: multiplier
};

struct WPHeader {
  int32_t p1C, p2C, p3[5], w[4];
  JXLB_HD void SetDefault() {
    p1C = 16;
    p2C = 10;
    p3[0] = p3[1] = p3[2] = 7;
    p3[3] = p3[4] = 0;
    w[0] = 13;
    w[1] = w[2] = w[3] = 12;
  }
};

static constexpr int kMaxModPlanes = 8;  // channels of a modular image once every transform is undone
struct ModTransform {
  uint8_t id;          // 0 RCT, 1 palette, 2 squeeze
  uint8_t rct_type;
  uint16_t pad;
  uint32_t begin_c;
  uint32_t num_c, nb_colours, nb_deltas, d_pred;
  // filled by PlanChannels: the output planes the transform works on (RCT: 3, palette: num_c; pl[0] also holds the
  // palette's index channel) and where the palette's colours sit in the stream's meta-channel buffer (int32 units)
  uint8_t pl[kMaxModPlanes];
  uint32_t meta_off;
};
static constexpr int kMaxTransforms = 8;

struct ModSqueezeParam {
  uint8_t horizontal, in_place;
  uint16_t pad;
  uint32_t begin_c, num_c;
};
static constexpr int kMaxSqueezeParams = 24;

struct ModularHeader {
  uint8_t use_global_tree;
  uint8_t nb_transforms;
  uint8_t has_squeeze;     // a squeeze transform is present (only accepted in a frame's global header, squeeze.h)
  uint8_t nb_squeeze;      // explicit parameter count (0: default parameters)
  ModSqueezeParam sq[kMaxSqueezeParams];
  WPHeader wp;
  ModTransform tr[kMaxTransforms];
};

JXLB_HD_NOINLINE int ReadModularHeader(BitReader& br, ModularHeader* h) {
  h->use_global_tree = (uint8_t) br.Read(1);
  h->wp.SetDefault();
  if (!br.Read(1)) {
    h->wp.p1C = (int32_t) br.Read(5);
    h->wp.p2C = (int32_t) br.Read(5);
    for (int i = 0; i < 5; ++i) h->wp.p3[i] = (int32_t) br.Read(5);
    for (int i = 0; i < 4; ++i) h->wp.w[i] = (int32_t) br.Read(4);
  }
  uint32_t nt = br.U32(0, 0, 1, 0, 2, 4, 18, 8);
  if (nt > (uint32_t) kMaxTransforms) return kErrUnsupported;
  h->nb_transforms = (uint8_t) nt;
  h->has_squeeze = 0;
  h->nb_squeeze = 0;
  for (uint32_t i = 0; i < nt; ++i) {
    ModTransform& t = h->tr[i];
    t.id = (uint8_t) br.Read(2);
    t.rct_type = 0;
    t.pad = 0;
    t.begin_c = t.num_c = t.nb_colours = t.nb_deltas = t.d_pred = 0;
    t.meta_off = 0;
    for (int k = 0; k < kMaxModPlanes; ++k) t.pl[k] = 0;
    if (t.id == 0) {
      t.begin_c = br.U32(0, 3, 8, 6, 72, 10, 1096, 13);
      t.rct_type = (uint8_t) br.U32(6, 0, 0, 2, 2, 4, 10, 6);
      if (t.rct_type >= 42) return kErrBadStream;
    } else if (t.id == 1) {
      t.begin_c = br.U32(0, 3, 8, 6, 72, 10, 1096, 13);
      t.num_c = br.U32(1, 0, 3, 0, 4, 0, 1, 13);
      t.nb_colours = br.U32(0, 8, 256, 10, 1280, 12, 5376, 16);
      t.nb_deltas = br.U32(0, 0, 1, 8, 257, 10, 1281, 16);
      t.d_pred = br.Read(4);
      // delta palettes: entries below nb_deltas, and the implicit colours of negative indices, are ADDED to a prediction
      // from the neighbouring output pixels (PaletteNeedsSerialInverse); libjxl's lossless encoder uses them now and then
      if (t.d_pred == 6 || t.d_pred > 13) return kErrUnsupported;  // a delta palette over the weighted predictor
      if (t.num_c > (uint32_t) kMaxModPlanes) return kErrUnsupported;
    } else if (t.id == 2) {
      if (h->has_squeeze) return kErrUnsupported;  // one squeeze transform per header
      h->has_squeeze = 1;
      const uint32_t nsq = br.U32(0, 0, 1, 4, 9, 6, 41, 8);
      if (nsq > (uint32_t) kMaxSqueezeParams) return kErrUnsupported;
      h->nb_squeeze = (uint8_t) nsq;
      for (uint32_t k = 0; k < nsq; ++k) {
        h->sq[k].horizontal = (uint8_t) br.Read(1);
        h->sq[k].in_place = (uint8_t) br.Read(1);
        h->sq[k].pad = 0;
        h->sq[k].begin_c = br.U32(0, 3, 8, 6, 72, 10, 1096, 13);
        h->sq[k].num_c = br.U32(1, 0, 2, 0, 3, 0, 4, 4);
      }
    } else {
      return kErrBadStream;
    }
  }
  return kOk;
}

// Decodes an MA tree into arena (TreeNode array).  Uses arena space after the tree temporarily for the tree code.
JXLB_HD_NOINLINE int DecodeTree(BitReader& br, Arena& arena, uint32_t max_nodes, uint32_t* tree_off, uint32_t* num_nodes,
                                uint32_t* uses_wp, uint32_t* max_property) {
  uint32_t toff = arena.Alloc(max_nodes * (uint32_t) sizeof(TreeNode), 16);
  if (toff == 0xFFFFFFFFu) return kErrScratch;
  uint32_t saved = arena.used;
  uint32_t coff;
  int st = ParseCode<false>(br, 6, true, arena, &coff);
  if (st != kOk) return st;
  CodeView cv;
  cv.Bind(arena.base + coff);
  uint32_t* win = nullptr;
  uint32_t wmask = 0;
  if (cv.lz77) {
    uint32_t wo = arena.Alloc(4u << 16, 16);  // tree streams are short; a 64K-entry window bounds max_nodes*6 symbols
    if (wo == 0xFFFFFFFFu) return kErrScratch;
    win = reinterpret_cast<uint32_t*>(arena.base + wo);
    wmask = (1u << 16) - 1;
    if (max_nodes * 6u > (1u << 16)) return kErrUnsupported;
  }
  SymbolReader sr;
  sr.Begin(cv, br, win, wmask);
  TreeNode* tree = reinterpret_cast<TreeNode*>(arena.base + toff);
  uint32_t n = 0, to_decode = 1, leaf = 0, wp = 0, maxp = 0;
  while (to_decode > 0) {
    if (n >= max_nodes) return kErrScratch;
    --to_decode;
    uint32_t p1 = ReadHybridUint(cv, sr, br, 1);
    if (p1 > 256) return kErrBadStream;
    TreeNode& nd = tree[n];
    nd.pad = 0;
    if (p1 == 0) {
      uint32_t pred = ReadHybridUint(cv, sr, br, 2);
      if (pred >= 14) return kErrBadStream;
      int32_t off = UnpackSigned(ReadHybridUint(cv, sr, br, 3));
      uint32_t mul_log = ReadHybridUint(cv, sr, br, 4);
      if (mul_log >= 31) return kErrBadStream;
      uint32_t mul_bits = ReadHybridUint(cv, sr, br, 5);
      if (mul_bits >= (1u << (31 - mul_log)) - 1) return kErrBadStream;
      nd.property = -1;
      nd.predictor = (uint8_t) pred;
      nd.split_or_offset = off;
      nd.left_or_ctx = leaf++;
      nd.right_or_mul = (mul_bits + 1) << mul_log;
      if (pred == 6) wp = 1;
    } else {
      int32_t sv = UnpackSigned(ReadHybridUint(cv, sr, br, 0));
      nd.property = (int16_t) (p1 - 1);
      nd.predictor = 0;
      nd.split_or_offset = sv;
      nd.left_or_ctx = n + to_decode + 1;
      nd.right_or_mul = n + to_decode + 2;
      to_decode += 2;
      if (p1 - 1 == 15) wp = 1;
      if (p1 - 1 > maxp) maxp = p1 - 1;
    }
    ++n;
  }
  if (!sr.FinalStateOk()) return kErrBadStream;
  arena.used = saved;  // drop the tree code; keep the nodes (over-allocated to max_nodes)
  *tree_off = toff;
  *num_nodes = n;
  *uses_wp = wp;
  *max_property = maxp;
  return kOk;
}

// One channel of a modular image as the decoder sees it.
struct ModChannel {
  int32_t* data;      // row-major int32 samples
  uint32_t w, h;
  uint32_t stride;    // in samples
};

// Weighted-predictor state (App. B.6); arrays live in the per-stream scratch.
struct WPState {
  int32_t* err;        // [2][w+2]
  int32_t* pred_err;   // [4][2][w+2]
  uint32_t xs;         // row length
  int64_t pred;        // last prediction (<<3)
  int64_t prediction[4];
  int32_t max_err_prop;
  JXLB_HD static uint32_t ScratchInts(uint32_t w) { return 10u * (w + 2u); }
  JXLB_HD void Init(int32_t* mem, uint32_t w) {
    xs = w;
    err = mem;
    pred_err = mem + 2 * (w + 2);
    for (uint32_t i = 0; i < 10u * (w + 2u); ++i) mem[i] = 0;
  }
};

JXLB_HD uint32_t WPDiv(uint32_t k) { return (1u << 24) / (k + 1); }

JXLB_HD int32_t WPErrWeight(uint32_t x, uint32_t maxweight) {
  int shift = FloorLog2(x + 1) - 5;
  if (shift < 0) shift = 0;
  return (int32_t) (4 + ((maxweight * WPDiv(x >> shift)) >> shift));
}

JXLB_HD int32_t WPPredict(WPState& s, const WPHeader& h, uint32_t x, uint32_t y, int32_t N_, int32_t W_, int32_t NE_,
                          int32_t NW_, int32_t NN_) {
  const uint32_t xs = s.xs;
  const uint32_t cur = (y & 1) ? 0 : xs + 2;
  const uint32_t prv = (y & 1) ? xs + 2 : 0;
  const uint32_t pN = prv + x;
  const uint32_t pNE = x < xs - 1 ? pN + 1 : pN;
  const uint32_t pNW = x > 0 ? pN - 1 : pN;
  uint32_t w[4];
  for (int i = 0; i < 4; ++i) {
    const int32_t* pe = s.pred_err + (size_t) i * 2 * (xs + 2);
    uint32_t e = (uint32_t) pe[pN] + (uint32_t) pe[pNE] + (uint32_t) pe[pNW];
    w[i] = (uint32_t) WPErrWeight(e, (uint32_t) h.w[i]);
  }
  int64_t N = (int64_t) N_ * 8, W = (int64_t) W_ * 8, NE = (int64_t) NE_ * 8, NW = (int64_t) NW_ * 8, NN = (int64_t) NN_ * 8;
  int64_t teW = x == 0 ? 0 : s.err[cur + x - 1];
  int64_t teN = s.err[pN];
  int64_t teNW = s.err[pNW];
  int64_t sumWN = teN + teW;
  int64_t teNE = s.err[pNE];
  int64_t p = teW;
  int64_t ap = p < 0 ? -p : p;
  if ((teN < 0 ? -teN : teN) > ap) { p = teN; ap = p < 0 ? -p : p; }
  if ((teNW < 0 ? -teNW : teNW) > ap) { p = teNW; ap = p < 0 ? -p : p; }
  if ((teNE < 0 ? -teNE : teNE) > ap) { p = teNE; }
  s.max_err_prop = (int32_t) p;
  s.prediction[0] = W + NE - N;
  s.prediction[1] = N - (((sumWN + teNE) * h.p1C) >> 5);
  s.prediction[2] = W - (((sumWN + teNW) * h.p2C) >> 5);
  s.prediction[3] = N - ((teNW * h.p3[0] + teN * h.p3[1] + teNE * h.p3[2] + (NN - N) * h.p3[3] + (NW - W) * h.p3[4]) >> 5);
  uint32_t ws = w[0] + w[1] + w[2] + w[3];
  int lw = FloorLog2(ws);
  for (int i = 0; i < 4; ++i) w[i] >>= (lw - 4);
  ws = w[0] + w[1] + w[2] + w[3];
  int64_t sum = (int64_t) (ws >> 1) - 1;
  for (int i = 0; i < 4; ++i) sum += s.prediction[i] * (int64_t) w[i];
  s.pred = (sum * (int64_t) WPDiv(ws - 1)) >> 24;
  if (((teN ^ teW) | (teN ^ teNW)) <= 0) {
    int64_t mx = W > NE ? W : NE;
    if (N > mx) mx = N;
    int64_t mn = W < NE ? W : NE;
    if (N < mn) mn = N;
    if (s.pred > mx) s.pred = mx;
    if (s.pred < mn) s.pred = mn;
  }
  return (int32_t) ((s.pred + 3) >> 3);
}

JXLB_HD void WPUpdate(WPState& s, int32_t val_, uint32_t x, uint32_t y) {
  const uint32_t xs = s.xs;
  const uint32_t cur = (y & 1) ? 0 : xs + 2;
  const uint32_t prv = (y & 1) ? xs + 2 : 0;
  int64_t val = (int64_t) val_ * 8;
  s.err[cur + x] = (int32_t) (s.pred - val);
  for (int i = 0; i < 4; ++i) {
    int32_t* pe = s.pred_err + (size_t) i * 2 * (xs + 2);
    int64_t d = s.prediction[i] - val;
    if (d < 0) d = -d;
    int32_t e = (int32_t) ((d + 3) >> 3);
    pe[cur + x] = e;
    pe[prv + x + 1] += e;
  }
}

JXLB_HD int32_t ClampedGradient(int32_t w, int32_t n, int32_t nw) {
  int32_t mn = w < n ? w : n, mx = w < n ? n : w;
  int64_t g = (int64_t) w + n - nw;
  return g < mn ? mn : g > mx ? mx : (int32_t) g;
}

// The 13 predictors that need no weighted-predictor state (ISO/IEC 18181-1 modular predictors; 6 = weighted is separate).
JXLB_HD int64_t PredictNoWp(uint32_t predictor, int32_t W, int32_t N, int32_t NW, int32_t NE, int32_t NN, int32_t WW, int32_t NEE) {
  switch (predictor) {
    case 0: return 0;
    case 1: return W;
    case 2: return N;
    case 3: return ((int64_t) W + N) / 2;
    case 4: {
      int64_t p = (int64_t) W + N - NW;
      int64_t pa = p - W, pb = p - N;
      if (pa < 0) pa = -pa;
      if (pb < 0) pb = -pb;
      return pa < pb ? W : N;
    }
    case 5: return ClampedGradient(W, N, NW);
    case 7: return NE;
    case 8: return NW;
    case 9: return WW;
    case 10: return ((int64_t) W + NW) / 2;
    case 11: return ((int64_t) N + NW) / 2;
    case 12: return ((int64_t) N + NE) / 2;
    default: return (6 * (int64_t) N - 2 * (int64_t) NN + 7 * (int64_t) W + WW + NEE + 3 * (int64_t) NE + 8) / 16;
  }
}

// Everything a stream needs besides the bits: tree + code (global or local) and scratch.
struct ModularContext {
  const TreeNode* tree;
  uint32_t num_nodes;
  uint32_t uses_wp;
  uint32_t max_property;
  CodeView code;
};

// Decodes channels [first, first+count) of `ch` (all channels listed share one sub-stream).  `ch` holds the whole
// channel list of the stream's image so that previous-channel properties can look back.  stream_id = property 1.
// scratch: WP state (WPState::ScratchInts(max w) ints) and, if the code uses LZ77, a window (see lz77_window).
JXLB_HD_NOINLINE int DecodeModularChannels(BitReader& br, const ModularContext& mc, const WPHeader& wph, const ModChannel* ch,
                                           uint32_t nch, uint32_t stream_id, int32_t* wp_scratch, uint32_t* lz77_window,
                                           uint32_t lz77_mask) {
  SymbolReader sr;
  uint32_t dist_mult = 0;
  for (uint32_t i = 0; i < nch; ++i)
    if (ch[i].w && ch[i].h && ch[i].w > dist_mult) dist_mult = ch[i].w;
  if (mc.code.lz77 && lz77_window == nullptr) return kErrUnsupported;
  sr.Begin(mc.code, br, lz77_window, lz77_mask);
  const TreeNode* tree = mc.tree;
  for (uint32_t ci = 0; ci < nch; ++ci) {
    const ModChannel& c = ch[ci];
    if (!c.w || !c.h) continue;
    WPState wps;
    if (mc.uses_wp) wps.Init(wp_scratch, c.w);
    for (uint32_t y = 0; y < c.h; ++y) {
      int32_t* row = c.data + (size_t) y * c.stride;
      const int32_t* rN = y > 0 ? row - c.stride : nullptr;
      const int32_t* rNN = y > 1 ? row - 2 * (size_t) c.stride : nullptr;
      int32_t prev9 = 0;
      for (uint32_t x = 0; x < c.w; ++x) {
        int32_t W = x > 0 ? row[x - 1] : (y > 0 ? rN[x] : 0);
        int32_t N = y > 0 ? rN[x] : W;
        int32_t NW = (x > 0 && y > 0) ? rN[x - 1] : W;
        int32_t NE = (x + 1 < c.w && y > 0) ? rN[x + 1] : N;
        int32_t NN = y > 1 ? rNN[x] : N;
        int32_t wp_pred = 0;
        if (mc.uses_wp) wp_pred = WPPredict(wps, wph, x, y, N, W, NE, NW, NN);
        int32_t p9 = (int32_t) ((int64_t) W + N - NW);
        // ---- tree walk
        uint32_t ni = 0;
        TreeNode nd = tree[0];
        while (nd.property >= 0) {
          int32_t v;
          switch (nd.property) {
            case 0: v = (int32_t) ci; break;
            case 1: v = (int32_t) stream_id; break;
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
            case 14: {
              int32_t WW = x > 1 ? row[x - 2] : W;
              v = W - WW;
              break;
            }
            case 15: v = wps.max_err_prop; break;
            default: {
              // previous-channel properties: 4 per earlier channel with identical dimensions, nearest first
              uint32_t k = (uint32_t) (nd.property - 16);
              uint32_t want = k >> 2, sub = k & 3, seen = 0;
              v = 0;
              for (int j = (int) ci - 1; j >= 0; --j) {
                if (ch[j].w != c.w || ch[j].h != c.h) continue;
                if (seen++ != want) continue;
                const int32_t* prow = ch[j].data + (size_t) y * ch[j].stride;
                int32_t rv = prow[x];
                if (sub == 0) v = rv < 0 ? -rv : rv;
                else if (sub == 1) v = rv;
                else {
                  int32_t vw = x > 0 ? prow[x - 1] : 0;
                  int32_t vn = y > 0 ? prow[(ptrdiff_t) x - (ptrdiff_t) ch[j].stride] : vw;
                  int32_t vnw = (x > 0 && y > 0) ? prow[(ptrdiff_t) x - 1 - (ptrdiff_t) ch[j].stride] : vw;
                  int32_t d = rv - ClampedGradient(vw, vn, vnw);
                  v = sub == 2 ? (d < 0 ? -d : d) : d;
                }
                break;
              }
              break;
            }
          }
          ni = v > nd.split_or_offset ? nd.left_or_ctx : nd.right_or_mul;
          nd = tree[ni];
        }
        prev9 = p9;
        // ---- prediction
        int64_t pred;
        switch (nd.predictor) {
          case 0: pred = 0; break;
          case 1: pred = W; break;
          case 2: pred = N; break;
          case 3: pred = ((int64_t) W + N) / 2; break;
          case 4: {
            int64_t p = (int64_t) W + N - NW;
            int64_t pa = p - W, pb = p - N;
            if (pa < 0) pa = -pa;
            if (pb < 0) pb = -pb;
            pred = pa < pb ? W : N;
            break;
          }
          case 5: pred = ClampedGradient(W, N, NW); break;
          case 6: pred = wp_pred; break;
          case 7: pred = NE; break;
          case 8: pred = NW; break;
          case 9: pred = x > 1 ? row[x - 2] : W; break;
          case 10: pred = ((int64_t) W + NW) / 2; break;
          case 11: pred = ((int64_t) N + NW) / 2; break;
          case 12: pred = ((int64_t) N + NE) / 2; break;
          default: {
            int64_t WW = x > 1 ? row[x - 2] : W;
            int64_t NEE = (x + 2 < c.w && y > 0) ? rN[x + 2] : NE;
            pred = (6 * (int64_t) N - 2 * (int64_t) NN + 7 * (int64_t) W + WW + NEE + 3 * (int64_t) NE + 8) / 16;
            break;
          }
        }
        uint32_t u = ReadHybridUint(mc.code, sr, br, nd.left_or_ctx, (int) dist_mult);
        int64_t val = (int64_t) UnpackSigned(u) * (int64_t) nd.right_or_mul + nd.split_or_offset + pred;
        int32_t out = (int32_t) val;
        row[x] = out;
        if (mc.uses_wp) WPUpdate(wps, out, x, y);
      }
    }
  }
  if (!sr.FinalStateOk()) return kErrBadStream;
  if (br.Overrun()) return kErrTruncated;
  return kOk;
}

// Inverse RCT on three int32 rows of length n (App. B.6).
JXLB_HD void InverseRctPixel(uint32_t type, int32_t& a, int32_t& b, int32_t& c) {
  uint32_t k = type % 7;
  int32_t F = a, S = b, T = c;
  if (k == 6) {
    int32_t tmp = F - (T >> 1);
    int32_t G = T + tmp;
    int32_t Bq = tmp - (S >> 1);
    int32_t Rq = Bq + S;
    F = Rq;
    S = G;
    T = Bq;
  } else {
    if (k & 1) T = T + F;
    if ((k >> 1) == 1) S = S + F;
    else if ((k >> 1) == 2) S = S + ((F + T) >> 1);
  }
  a = F;
  b = S;
  c = T;
}
// Output channel permutation of RCT `type`: first/second/third results go to begin_c + out[0..2].
JXLB_HD void RctPermutation(uint32_t type, uint32_t out[3]) {
  uint32_t perm = type / 7;
  out[0] = perm % 3;
  out[1] = (perm + 1 + perm / 3) % 3;
  out[2] = (perm + 2 - perm / 3) % 3;
}

// ---- channel list of a stream whose header carries transforms ---------------------------------------------------------
// A palette transform replaces num_c channels by one index channel and puts a "meta" channel (nb_colours x num_c, the
// colours) at the front of the stream's channel list; begin_c of every later transform counts those meta channels
// (ISO/IEC 18181-1 modular transforms; libjxl's meta-apply step).  PlanChannels replays the header's transforms on the
// list of the nfinal output planes and records, per transform, the planes it works on, so that the inverse can run per
// pixel on planes in place: the index channel of a palette lives in the first of its output planes.
struct ChannelPlan {
  uint32_t nb_meta;                      // meta channels, first in the stream
  uint32_t ncoded;                       // channels that follow them
  uint8_t coded_plane[kMaxModPlanes];    // output plane holding coded channel i
  uint8_t meta_tr[kMaxTransforms];       // transform owning meta channel i
  uint32_t meta_ints;                    // size of the meta-channel buffer
};

// Entries of a palette's meta channel per colour channel: nb_deltas delta entries (the indices below nb_deltas) followed
// by nb_colours colours.
JXLB_HD uint32_t PaletteWidth(const ModTransform& tr) { return tr.nb_colours + tr.nb_deltas; }
// True when the inverse has to walk the pixels in raster order: a value may be added to a prediction from its neighbours.
JXLB_HD bool PaletteNeedsSerialInverse(const ModTransform& tr) { return tr.id == 1 && (tr.nb_deltas != 0 || tr.d_pred != 0); }

JXLB_HD int PlanChannels(ModularHeader* mh, uint32_t nfinal, ChannelPlan* cp) {
  if (nfinal > (uint32_t) kMaxModPlanes) return kErrUnsupported;
  uint8_t list[kMaxModPlanes + kMaxTransforms];  // plane id, or 0x80 | transform for a meta channel
  uint32_t len = nfinal, nb_meta = 0;
  for (uint32_t i = 0; i < nfinal; ++i) list[i] = (uint8_t) i;
  cp->meta_ints = 0;
  for (uint32_t t = 0; t < mh->nb_transforms; ++t) {
    ModTransform& tr = mh->tr[t];
    if (tr.id == 0) {
      if (tr.begin_c + 3 > len) return kErrBadStream;
      if (tr.begin_c < nb_meta) return kErrUnsupported;  // RCT across palette colours
      for (int k = 0; k < 3; ++k) tr.pl[k] = list[tr.begin_c + k];
    } else if (tr.id == 1) {
      if (tr.num_c == 0 || tr.begin_c + tr.num_c > len) return kErrBadStream;
      if (tr.begin_c < nb_meta) return kErrUnsupported;  // palette of a palette
      if (tr.nb_colours > (1u << 20) || tr.nb_deltas > (1u << 20) ||
          cp->meta_ints + (uint64_t) PaletteWidth(tr) * tr.num_c > (16u << 20))
        return kErrUnsupported;
      for (uint32_t k = 0; k < tr.num_c; ++k) tr.pl[k] = list[tr.begin_c + k];
      for (uint32_t i = tr.begin_c + tr.num_c; i < len; ++i) list[i - (tr.num_c - 1)] = list[i];
      len -= tr.num_c - 1;
      for (uint32_t i = len; i > 0; --i) list[i] = list[i - 1];
      list[0] = (uint8_t) (0x80u | t);
      ++len;
      ++nb_meta;
      tr.meta_off = cp->meta_ints;
      cp->meta_ints += PaletteWidth(tr) * tr.num_c;
    } else {
      return kErrUnsupported;  // squeeze: handled by its own path (squeeze.h), never together with the others here
    }
  }
  cp->nb_meta = nb_meta;
  cp->ncoded = len - nb_meta;
  for (uint32_t i = 0; i < nb_meta; ++i) cp->meta_tr[i] = (uint8_t) (list[i] & 0x7Fu);
  for (uint32_t i = nb_meta; i < len; ++i) cp->coded_plane[i - nb_meta] = list[i];
  return kOk;
}

// Colours outside the coded palette (ISO/IEC 18181-1: negative indices address a fixed table of 72 signed deltas,
// indices past the palette two implicit colour cubes).  Table as shipped in the reference's libjxl 0.12.0.
JXLB_HD_NOINLINE int32_t ImplicitPaletteValue(int32_t index, uint32_t c, int32_t palette_size, uint32_t bit_depth) {
  if (c >= 3) return 0;
  if (index < 0) {
    const int16_t kDelta[72 * 3] = {
        0, 0, 0, 4, 4, 4, 11, 0, 0, 0, 0, -13, 0, -12, 0, -10, -10, -10, -18, -18, -18, -27, -27, -27, -18, -18, 0, 0, 0, -32, -32, 0, 0, -37, -37,
        -37, 0, -32, -32, 24, 24, 45, 50, 50, 50, -45, -24, -24, -24, -45, -45, 0, -24, -24, -34, -34, 0, -24, 0, -24, -45, -45, -24, 64, 64, 64,
        -32, 0, -32, 0, -32, 0, -32, 0, 32, -24, -45, -24, 45, 24, 45, 24, -24, -45, -45, -24, 24, 80, 80, 80, 64, 0, 0, 0, 0, -64, 0, -64, -64, -24,
        -24, 45, 96, 96, 96, 64, 64, 0, 45, -24, -24, 34, -34, 0, 112, 112, 112, 24, -45, -45, 45, 45, -24, 0, -32, 32, 24, -24, 45, 0, 96, 96, 45,
        -24, 24, 24, -45, -24, -24, -45, 24, 0, -64, 0, 96, 0, 0, 128, 128, 128, 64, 0, 64, 144, 144, 144, 96, 96, 0, -36, -36, 36, 45, -24, -45, 45,
        -45, -24, 0, 0, -96, 0, 128, 128, 0, 96, 0, 45, 24, -45, -128, 0, 0, 24, -45, 24, -45, 24, -45, 64, 0, -64, 64, -64, -64, 96, 0, 96, 45, -45,
        24, 24, 45, -45, 64, 64, -64, 128, 128, 0, 0, 0, -128, -24, 45, -45};
    int32_t i = -(index + 1);
    i %= 143;  // 1 + 2 * (72 - 1)
    const int32_t e = (i + 1) >> 1;
    int32_t v = kDelta[e * 3 + c];
    v *= (i & 1) ? 1 : -1;
    if (bit_depth > 8) v *= 1 << (bit_depth - 8);
    return v;
  }
  index -= palette_size;
  if (index < 64) {  // small cube: 4 x 4 x 4
    index >>= 2 * c;
    return (int32_t) (((int64_t) (index % 4) * ((1ll << bit_depth) - 1)) / 4 + (1ll << (bit_depth > 3 ? bit_depth - 3 : 0)));
  }
  index -= 64;       // large cube: 5 x 5 x 5
  if (c == 1) index /= 5;
  else if (c == 2) index /= 25;
  return (int32_t) (((int64_t) (index % 5) * ((1ll << bit_depth) - 1)) / 4);
}

JXLB_HD int32_t PaletteValue(const int32_t* pal, int32_t index, uint32_t c, uint32_t width, uint32_t bit_depth) {
  if (index >= 0 && (uint32_t) index < width) return pal[(size_t) c * width + (uint32_t) index];
  return ImplicitPaletteValue(index, c, (int32_t) width, bit_depth);
}

// One pixel of the inverse palette: the index sits in the first output plane.  False: a negative index (an implicit delta
// colour) under a predictor, which would need the neighbouring output pixels -- not covered.
JXLB_HD bool InversePalettePixel(const ModTransform& tr, const int32_t* meta, uint32_t bit_depth, int32_t index, int32_t* out) {
  const int32_t* pal = meta + tr.meta_off;
  const bool ok = !(PaletteNeedsSerialInverse(tr) && index < (int32_t) tr.nb_deltas);  // callers route such palettes to the serial inverse
  if (tr.num_c == 1) {  // single-channel palettes clamp the index instead of using implicit colours
    const int32_t hi = (int32_t) PaletteWidth(tr) - 1;
    if (index > hi) index = hi;
    if (index < 0) index = hi < 0 ? -1 : 0;
  }
  for (uint32_t c = 0; c < tr.num_c; ++c) out[c] = PaletteValue(pal, index, c, PaletteWidth(tr), bit_depth);
  return ok;
}

// Serial inverse of transforms tr[0 .. n) (undone last to first) over whole planes (`planes`: the stream's output planes, all
// of one size).  The per-pixel form for frame-level transforms of large images is StageGlobalInverse (pixel_stages.h);
// palettes with deltas or a predictor only exist in this form: their pixels are visited in raster order and a delta is
// added to the prediction from the already reconstructed neighbours of the same output plane.
JXLB_HD int ApplyInverseTransforms(const ModTransform* trs, uint32_t n, const ModChannel* planes, const int32_t* meta, uint32_t bit_depth) {
  bool ok = true;
  for (int t = (int) n - 1; t >= 0; --t) {
    const ModTransform& tr = trs[t];
    if (tr.id == 0) {
      const ModChannel& a = planes[tr.pl[0]];
      const ModChannel& b = planes[tr.pl[1]];
      const ModChannel& c = planes[tr.pl[2]];
      uint32_t perm[3];
      RctPermutation(tr.rct_type, perm);
      const ModChannel* dst[3] = {&planes[tr.pl[perm[0]]], &planes[tr.pl[perm[1]]], &planes[tr.pl[perm[2]]]};
      for (uint32_t y = 0; y < a.h; ++y) {
        for (uint32_t x = 0; x < a.w; ++x) {
          int32_t v0 = a.data[(size_t) y * a.stride + x];
          int32_t v1 = b.data[(size_t) y * b.stride + x];
          int32_t v2 = c.data[(size_t) y * c.stride + x];
          InverseRctPixel(tr.rct_type, v0, v1, v2);
          dst[0]->data[(size_t) y * dst[0]->stride + x] = v0;
          dst[1]->data[(size_t) y * dst[1]->stride + x] = v1;
          dst[2]->data[(size_t) y * dst[2]->stride + x] = v2;
        }
      }
    } else if (tr.id == 1 && !PaletteNeedsSerialInverse(tr)) {
      const ModChannel& ic = planes[tr.pl[0]];
      for (uint32_t y = 0; y < ic.h; ++y) {
        for (uint32_t x = 0; x < ic.w; ++x) {
          int32_t v[kMaxModPlanes];
          ok &= InversePalettePixel(tr, meta, bit_depth, ic.data[(size_t) y * ic.stride + x], v);
          for (uint32_t c = 0; c < tr.num_c; ++c) planes[tr.pl[c]].data[(size_t) y * planes[tr.pl[c]].stride + x] = v[c];
        }
      }
    } else if (tr.id == 1) {
      // delta palette.  The index plane is output plane 0 of the transform; at (x, y) every neighbour a predictor looks at
      // (W, WW, N, NW, NE, NEE, NN) has already been turned into an output sample, in every plane.
      const ModChannel& ic = planes[tr.pl[0]];
      const int32_t* pal = meta + tr.meta_off;
      const uint32_t width = PaletteWidth(tr);
      for (uint32_t y = 0; y < ic.h; ++y) {
        for (uint32_t x = 0; x < ic.w; ++x) {
          const int32_t index = ic.data[(size_t) y * ic.stride + x];
          for (uint32_t c = 0; c < tr.num_c; ++c) {
            const ModChannel& pc = planes[tr.pl[c]];
            int32_t* pp = pc.data + (size_t) y * pc.stride + x;
            int64_t val = PaletteValue(pal, index, c, width, bit_depth);
            if (index < (int32_t) tr.nb_deltas) {
              const ptrdiff_t row = (ptrdiff_t) pc.stride;
              const int32_t W = x ? pp[-1] : (y ? pp[-row] : 0);
              const int32_t N = y ? pp[-row] : W;
              const int32_t NW = (x && y) ? pp[-1 - row] : W;
              const int32_t NE = (x + 1 < pc.w && y) ? pp[1 - row] : N;
              const int32_t WW = x > 1 ? pp[-2] : W;
              const int32_t NN = y > 1 ? pp[-2 * row] : N;
              const int32_t NEE = (x + 2 < pc.w && y) ? pp[2 - row] : NE;
              val += PredictNoWp(tr.d_pred, W, N, NW, NE, NN, WW, NEE);
            }
            *pp = (int32_t) val;
          }
        }
      }
    }
  }
  return ok ? kOk : kErrUnsupported;
}

}  // namespace jxlb
