// Latency-optimised modular channel decoder (host + device), functionally identical to DecodeModularChannels in
// modular.h.  A modular sub-stream is one serial dependency chain (ANS state -> symbol -> sample -> next context), so
// on the GPU its speed is set by the latency of that chain in ONE lane.  This version therefore:
//   * keeps the bit reader, ANS state and code view in registers (local copies; no aliasing with the sample stores);
//   * prunes the MA tree per channel by the static properties (channel index, stream id) and classifies the rest:
//     - "WP-only" trees (every test on the weighted-predictor error, every leaf Weighted/offset 0/multiplier 1 — what the
//       streaming encoder emits for LF coefficients) become a 1024-entry context LUT instead of a tree walk,
//     - single-leaf subtrees skip the walk altogether,
//     - the weighted predictor state is only maintained when the channel's subtree needs it;
//   * slides the N/NW/NE neighbourhood of samples and of the weighted-predictor error rows through registers, reading
//     one new value per array per sample from small row buffers in caller-provided scratch (shared memory on device)
//     and never re-reading what it just wrote.
// Format digest: SURVEY.md App. B.6.  Replaces, for this path, libjxl 0.12.0's modular decoder behind the reference's
// DecodeJpegXlOneShot (/root/reference/jxlcoder/src/main/cpp/interop/JxlDecoding.cpp:74-175).
#pragma once
#include "modular.h"

namespace jxlb {

// Scratch layout (int32 units) for channels up to `w` wide.
struct ModFastScratch {
  static constexpr uint32_t kLutEntries = 1024;
  JXLB_HD static uint32_t Ints(uint32_t w) { return 64 + kLutEntries / 2 + 3 * (w + 8) + 10 * (w + 2); }
};

// Register-resident weighted-predictor state of the current sample position.
struct WPRegs {
  int32_t pe_nw[4], pe_n[4];  // predictor error sums row above at x-1 and x (with the running "+= e" applied)
  int32_t err_nw, err_n, err_w;
};

// One weighted-predictor step: prediction (<<3 domain handled inside), returns the prediction and the max-error
// property; `pe_ne` / `err_ne`: values of the row above at x+1 (or x at the right edge).
// (32-bit arithmetic: sample magnitudes up to 2^18 keep every intermediate below 2^31; only the final
// multiplication by the 24-bit reciprocal needs 64 bits.)
struct WPOut {
  int32_t pred;           // clamped weighted prediction, <<3
  int32_t sub[4];         // the four sub-predictions, <<3
  int32_t prop;           // property 15
};

JXLB_HD void WPStep(const WPRegs& r, const int32_t pe_ne[4], int32_t err_ne, const WPHeader& h, const uint32_t* divtab,
                    int32_t N_, int32_t W_, int32_t NE_, int32_t NW_, int32_t NN_, WPOut* o) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t e = (uint32_t) r.pe_n[i] + (uint32_t) pe_ne[i] + (uint32_t) r.pe_nw[i];
    int shift = FloorLog2(e + 1) - 5;
    if (shift < 0) shift = 0;
    w[i] = 4 + (((uint32_t) h.w[i] * divtab[e >> shift]) >> shift);
  }
  const int32_t N = N_ * 8, W = W_ * 8, NE = NE_ * 8, NW = NW_ * 8, NN = NN_ * 8;
  const int32_t teW = r.err_w, teN = r.err_n, teNW = r.err_nw, teNE = err_ne;
  const int32_t sumWN = teN + teW;
  int32_t p = teW;
  int32_t ap = p < 0 ? -p : p;
  if ((teN < 0 ? -teN : teN) > ap) { p = teN; ap = p < 0 ? -p : p; }
  if ((teNW < 0 ? -teNW : teNW) > ap) { p = teNW; ap = p < 0 ? -p : p; }
  if ((teNE < 0 ? -teNE : teNE) > ap) { p = teNE; }
  o->prop = (int32_t) p;
  o->sub[0] = W + NE - N;
  o->sub[1] = N - (((sumWN + teNE) * h.p1C) >> 5);
  o->sub[2] = W - (((sumWN + teNW) * h.p2C) >> 5);
  o->sub[3] = N - ((teNW * h.p3[0] + teN * h.p3[1] + teNE * h.p3[2] + (NN - N) * h.p3[3] + (NW - W) * h.p3[4]) >> 5);
  uint32_t ws = w[0] + w[1] + w[2] + w[3];
  const int lw = FloorLog2(ws);
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] >>= (lw - 4);
  ws = w[0] + w[1] + w[2] + w[3];
  int32_t sum = (int32_t) (ws >> 1) - 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) sum += o->sub[i] * (int32_t) w[i];
  int32_t pred = (int32_t) (((int64_t) sum * (int64_t) divtab[ws - 1]) >> 24);
  if (((teN ^ teW) | (teN ^ teNW)) <= 0) {
    int32_t mx = W > NE ? W : NE;
    if (N > mx) mx = N;
    int32_t mn = W < NE ? W : NE;
    if (N < mn) mn = N;
    if (pred > mx) pred = mx;
    if (pred < mn) pred = mn;
  }
  o->pred = pred;
}

// Classification of the part of the tree a channel can reach.
struct SubtreeInfo {
  uint32_t root;
  bool single_leaf, wp_only, uses_wp, needs_general;
};

JXLB_HD_NOINLINE SubtreeInfo AnalyseSubtree(const TreeNode* tree, uint32_t num_nodes, uint32_t ci, uint32_t stream_id) {
  SubtreeInfo s;
  uint32_t root = 0;
  for (uint32_t guard = 0; guard < num_nodes; ++guard) {
    const TreeNode nd = tree[root];
    if (nd.property == 0) root = (int32_t) ci > nd.split_or_offset ? nd.left_or_ctx : nd.right_or_mul;
    else if (nd.property == 1) root = (int32_t) stream_id > nd.split_or_offset ? nd.left_or_ctx : nd.right_or_mul;
    else break;
  }
  s.root = root;
  s.single_leaf = tree[root].property < 0;
  s.wp_only = true;
  s.uses_wp = false;
  s.needs_general = false;
  // iterative DFS over the reachable nodes (children always have larger indices than their parent)
  uint32_t stack[64];
  int sp = 0;
  stack[sp++] = root;
  uint32_t visited = 0;
  while (sp > 0) {
    const TreeNode nd = tree[stack[--sp]];
    if (++visited > num_nodes) break;
    if (nd.property < 0) {
      if (nd.predictor == 6) s.uses_wp = true;
      if (nd.predictor != 6 || nd.split_or_offset != 0 || nd.right_or_mul != 1) s.wp_only = false;
    } else {
      if (nd.property == 15) s.uses_wp = true;
      if (nd.property != 15 || nd.split_or_offset < -511 || nd.split_or_offset > 510) s.wp_only = false;
      if (sp + 2 > 64) {  // pathological depth: give up on the fast paths
        s.wp_only = false;
        s.uses_wp = true;
        s.needs_general = true;
        break;
      }
      stack[sp++] = nd.left_or_ctx;
      stack[sp++] = nd.right_or_mul;
    }
  }
  if (s.single_leaf) s.wp_only = false;
  return s;
}

JXLB_HD_NOINLINE int DecodeModularChannelsFast(BitReader& br_io, const ModularContext& mc, const WPHeader& wph, const ModChannel* ch,
                                               uint32_t nch, uint32_t stream_id, int32_t* scratch, uint32_t* lz77_window,
                                               uint32_t lz77_mask, int32_t* fast_scratch = nullptr, uint32_t fast_ints = 0) {
  BitReader br = br_io;       // register copies; written back on exit
  const CodeView code = mc.code;
  SymbolReader sr;
  uint32_t dist_mult = 0, maxw = 0;
  for (uint32_t i = 0; i < nch; ++i)
    if (ch[i].w && ch[i].h) {
      if (ch[i].w > dist_mult) dist_mult = ch[i].w;
      if (ch[i].w > maxw) maxw = ch[i].w;
    }
  if (code.lz77 && lz77_window == nullptr) return kErrUnsupported;
  sr.Begin(code, br, lz77_window, lz77_mask);
  const TreeNode* tree = mc.tree;
  int status = kOk;
  (void) maxw;
  for (uint32_t ci = 0; ci < nch && status == kOk; ++ci) {
    const ModChannel c = ch[ci];
    if (!c.w || !c.h) continue;
    const SubtreeInfo info = AnalyseSubtree(tree, mc.num_nodes, ci, stream_id);
    const uint32_t xs = c.w;
    // scratch carve-up for this channel: the fast (shared-memory) region when the channel is narrow enough
    int32_t* base = (fast_scratch && ModFastScratch::Ints(xs) <= fast_ints) ? fast_scratch : scratch;
    uint32_t* divtab = reinterpret_cast<uint32_t*>(base);
    uint16_t* lut = reinterpret_cast<uint16_t*>(base + 64);
    int32_t* rows = base + 64 + ModFastScratch::kLutEntries / 2;
    int32_t* wpmem = rows + 3 * (xs + 8);
    for (uint32_t k = 0; k < 64; ++k) divtab[k] = (1u << 24) / (k + 1);
    int32_t* rbuf[3] = {rows, rows + (xs + 8), rows + 2 * (xs + 8)};  // rotating: [0] current, [1] N row, [2] NN row
    // weighted-predictor rows: err[2][xs+2], pe[4][2][xs+2]
    int32_t* err_rows = wpmem;
    int32_t* pe_rows = wpmem + 2 * (xs + 2);
    if (info.uses_wp)
      for (uint32_t k = 0; k < 10u * (xs + 2u); ++k) wpmem[k] = 0;
    if (info.wp_only) {
      // context LUT over the clamped property value
      for (int v = -512; v < 512; ++v) {
        uint32_t ni = info.root;
        TreeNode nd = tree[ni];
        while (nd.property >= 0) {
          ni = v > nd.split_or_offset ? nd.left_or_ctx : nd.right_or_mul;
          nd = tree[ni];
        }
        lut[v + 512] = (uint16_t) nd.left_or_ctx;
      }
    }
    const TreeNode root_node = tree[info.root];
    // NN feeds WP sub-predictor 3 only through p3[3], and otherwise only general trees (property 13 / predictor 13)
    const bool need_nn = !info.wp_only || wph.p3[3] != 0;
    for (uint32_t y = 0; y < c.h; ++y) {
      int32_t* out_row = c.data + (size_t) y * c.stride;
      int32_t* cur = rbuf[0];
      const int32_t* rN = rbuf[1];
      const int32_t* rNN = rbuf[2];
      const uint32_t cur_o = (y & 1) ? 0 : xs + 2, prv_o = (y & 1) ? xs + 2 : 0;
      WPRegs wr;
      if (info.uses_wp) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int32_t* pe = pe_rows + (size_t) i * 2 * (xs + 2) + prv_o;
          wr.pe_n[i] = pe[0];
          wr.pe_nw[i] = pe[0];  // x == 0: NW position == N position
        }
        wr.err_n = err_rows[prv_o];
        wr.err_nw = wr.err_n;
        wr.err_w = 0;
      }
      // sliding sample neighbourhood
      int32_t W = y > 0 ? rN[0] : 0;
      int32_t N = y > 0 ? rN[0] : W;
      int32_t NW = W;
      int32_t WW = W;
      int32_t prev9 = 0;
      for (uint32_t x = 0; x < xs; ++x) {
        const int32_t NE = (x + 1 < xs && y > 0) ? rN[x + 1] : N;
        const int32_t NN = (need_nn && y > 1) ? rNN[x] : N;
        WPOut wo;
        int32_t pe_ne[4];
        int32_t err_ne = 0;
        if (info.uses_wp) {
          const bool edge = !(x + 1 < xs);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int32_t* pe = pe_rows + (size_t) i * 2 * (xs + 2) + prv_o;
            pe_ne[i] = edge ? wr.pe_n[i] : pe[x + 1];
          }
          err_ne = edge ? wr.err_n : err_rows[prv_o + x + 1];
          WPStep(wr, pe_ne, err_ne, wph, divtab, N, W, NE, NW, NN, &wo);
        }
        uint32_t ctx, predictor, mul;
        int32_t offset;
        if (info.wp_only) {
          int v = wo.prop;
          v = v < -512 ? -512 : v > 511 ? 511 : v;
          ctx = lut[v + 512];
          predictor = 6;
          offset = 0;
          mul = 1;
        } else {
          TreeNode nd = root_node;
          const int32_t p9 = (int32_t) ((int64_t) W + N - NW);
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
              case 14: v = W - WW; break;
              case 15: v = wo.prop; break;
              default: {
                const uint32_t k = (uint32_t) (nd.property - 16);
                const uint32_t want = k >> 2, sub = k & 3;
                uint32_t seen = 0;
                v = 0;
                for (int j = (int) ci - 1; j >= 0; --j) {
                  if (ch[j].w != c.w || ch[j].h != c.h) continue;
                  if (seen++ != want) continue;
                  const int32_t* prow = ch[j].data + (size_t) y * ch[j].stride;
                  const int32_t rv = prow[x];
                  if (sub == 0) v = rv < 0 ? -rv : rv;
                  else if (sub == 1) v = rv;
                  else {
                    const int32_t vw = x > 0 ? prow[x - 1] : 0;
                    const int32_t vn = y > 0 ? prow[(ptrdiff_t) x - (ptrdiff_t) ch[j].stride] : vw;
                    const int32_t vnw = (x > 0 && y > 0) ? prow[(ptrdiff_t) x - 1 - (ptrdiff_t) ch[j].stride] : vw;
                    const int32_t d = rv - ClampedGradient(vw, vn, vnw);
                    v = sub == 2 ? (d < 0 ? -d : d) : d;
                  }
                  break;
                }
                break;
              }
            }
            nd = tree[v > nd.split_or_offset ? nd.left_or_ctx : nd.right_or_mul];
          }
          prev9 = p9;
          ctx = nd.left_or_ctx;
          predictor = nd.predictor;
          offset = nd.split_or_offset;
          mul = nd.right_or_mul;
        }
        int64_t pred;
        if (predictor == 6) {
          pred = (wo.pred + 3) >> 3;
        } else {
          const int32_t NEE = (x + 2 < xs && y > 0) ? rN[x + 2] : NE;
          pred = PredictNoWp(predictor, W, N, NW, NE, NN, WW, NEE);
        }
        const uint32_t u = ReadHybridUint(code, sr, br, ctx, (int) dist_mult);
        const int32_t val = (int32_t) ((int64_t) UnpackSigned(u) * (int64_t) mul + offset + pred);
        cur[x] = val;
        out_row[x] = val;
        if (info.uses_wp) {
          const int32_t v8 = val * 8;
          const int32_t e_cur = wo.pred - v8;
          err_rows[cur_o + x] = e_cur;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            int32_t d = wo.sub[i] - v8;
            if (d < 0) d = -d;
            const int32_t e = (d + 3) >> 3;
            pe_rows[(size_t) i * 2 * (xs + 2) + cur_o + x] = e;
            wr.pe_nw[i] = wr.pe_n[i];
            wr.pe_n[i] = pe_ne[i] + e;  // the row above at x+1, plus this sample's "+= e"
          }
          wr.err_nw = wr.err_n;
          wr.err_n = err_ne;
          wr.err_w = e_cur;
        }
        // slide the sample neighbourhood (first row: N and NW fall back to W; x == 1: WW falls back to W)
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
      // rotate row buffers: current becomes N row, N row becomes NN row
      int32_t* t = rbuf[2];
      rbuf[2] = rbuf[1];
      rbuf[1] = rbuf[0];
      rbuf[0] = t;
    }
  }
  br_io = br;
  if (status != kOk) return status;
  if (!sr.FinalStateOk()) return kErrBadStream;
  if (br.Overrun()) return kErrTruncated;
  return kOk;
}

}  // namespace jxlb
