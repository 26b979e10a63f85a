// Squeeze (the Haar-like integer wavelet of JPEG XL's modular mode), as libjxl 0.12.0 applies it to lossy extra
// channels: with distance > 0 the alpha channel of a VarDCT frame is coded as a pyramid of residual channels.
// The reference reaches this through DecodeJpegXlOneShot (/root/reference/jxlcoder/src/main/cpp/interop/JxlDecoding.cpp:
// 36-176); 3 of the 13 files its demo app ships use it.  Restated from ISO/IEC 18181-1 (squeeze transform, default
// parameters, smooth tendency) and pinned end-to-end against the reference binary (tests/test_squeeze_host.py).
//
// Host side: SqueezeLayout() replays the forward "meta" transform on the channel list (which channels exist, their sizes
// and shifts, in bitstream order) and records the inverse steps.  Device / host: InvSqueezeRow / InvSqueezeColumn undo one
// step for one row / column -- the horizontal step is serial along x (the tendency looks at the pixel just
// reconstructed), the vertical one along y, so a step is parallel over rows or columns only.
#pragma once
#include "hd.h"

#include <vector>

namespace jxlb {

struct SqChannel {
  uint32_t w, h;
  uint32_t hshift, vshift;
  uint32_t off;   // int32 offset of the channel's samples in the squeeze buffer (row stride = w)
  uint32_t pad;
};

struct SqStep {   // out = unsqueeze(avg, res)
  uint32_t horizontal;
  uint32_t avg_off, res_off, out_off;       // int32 offsets in the squeeze buffer; out_off = 0xFFFFFFFF: the channel's final plane
  uint32_t avg_w, avg_h, res_w, res_h;      // out is (avg_w + res_w) x avg_h (horizontal) or avg_w x (avg_h + res_h)
  uint32_t final_channel;                   // which channel of the modular image the final plane is
  uint32_t pad;
};

JXLB_HD int32_t SmoothTendency(int32_t B, int32_t a, int32_t n) {
  int32_t diff = 0;
  if (B >= a && a >= n) {
    diff = (4 * B - 3 * n - a + 6) / 12;
    if (diff - (diff & 1) > 2 * (B - a)) diff = 2 * (B - a) + 1;
    if (diff + (diff & 1) > 2 * (a - n)) diff = 2 * (a - n);
  } else if (B <= a && a <= n) {
    diff = (4 * B - 3 * n - a - 6) / 12;
    if (diff + (diff & 1) < 2 * (B - a)) diff = 2 * (B - a) - 1;
    if (diff - (diff & 1) < 2 * (a - n)) diff = 2 * (a - n);
  }
  return diff;
}

// One row of a horizontal step: avg[0..aw), res[0..rw) -> out[0..aw+rw)
JXLB_HD void InvSqueezeRow(const int32_t* avg, const int32_t* res, int32_t* out, uint32_t aw, uint32_t rw) {
  int32_t left = 0;
  for (uint32_t x = 0; x < rw; ++x) {
    const int32_t a = avg[x];
    const int32_t next = x + 1 < aw ? avg[x + 1] : a;
    if (x == 0) left = a;
    const int32_t diff = res[x] + SmoothTendency(left, a, next);
    const int32_t A = a + diff / 2;
    out[2 * x] = A;
    left = A - diff;
    out[2 * x + 1] = left;
  }
  if (aw > rw) out[2 * rw] = avg[rw];
}

// One column of a vertical step; strides in samples
JXLB_HD void InvSqueezeColumn(const int32_t* avg, uint32_t avg_stride, const int32_t* res, uint32_t res_stride, int32_t* out,
                              uint32_t out_stride, uint32_t ah, uint32_t rh) {
  int32_t top = 0;
  for (uint32_t y = 0; y < rh; ++y) {
    const int32_t a = avg[(size_t) y * avg_stride];
    const int32_t next = y + 1 < ah ? avg[(size_t) (y + 1) * avg_stride] : a;
    if (y == 0) top = a;
    const int32_t diff = res[(size_t) y * res_stride] + SmoothTendency(top, a, next);
    const int32_t A = a + diff / 2;
    out[(size_t) (2 * y) * out_stride] = A;
    top = A - diff;
    out[(size_t) (2 * y + 1) * out_stride] = top;
  }
  if (ah > rh) out[(size_t) (2 * rh) * out_stride] = avg[(size_t) rh * avg_stride];
}

struct SqueezeParam {
  bool horizontal, in_place;
  uint32_t begin_c, num_c;
};

// Default parameters (no explicit list coded) for nch channels of w x h.
inline std::vector<SqueezeParam> DefaultSqueezeParams(uint32_t nch, uint32_t w, uint32_t h) {
  std::vector<SqueezeParam> ps;
  if (nch > 2) {  // channels 1 and 2 treated as chroma first (same sizes here by construction)
    ps.push_back({true, false, 1, 2});
    ps.push_back({false, false, 1, 2});
  }
  SqueezeParam p{true, true, 0, nch};
  const bool wide = w > h;
  if (!wide && h > 8) {
    p.horizontal = false;
    ps.push_back(p);
    h = (h + 1) / 2;
  }
  while (w > 8 || h > 8) {
    if (w > 8) {
      p.horizontal = true;
      ps.push_back(p);
      w = (w + 1) / 2;
    }
    if (h > 8) {
      p.horizontal = false;
      ps.push_back(p);
      h = (h + 1) / 2;
    }
  }
  return ps;
}

struct SqueezeLayoutOut {
  std::vector<SqChannel> channels;  // in bitstream order
  std::vector<SqStep> steps;        // in the order they must be applied
  size_t buffer_ints = 0;           // squeeze buffer size
};

// nch same-size channels (w x h); final planes are reported through SqStep::out_off = 0xFFFFFFFF + final_channel.
inline bool SqueezeLayout(const std::vector<SqueezeParam>& params, uint32_t nch, uint32_t w, uint32_t h, SqueezeLayoutOut* out) {
  struct Node {
    uint32_t w, h, hs, vs, off;
    int final_channel;   // >= 0: this node is original channel k (its buffer is the channel's plane)
  };
  std::vector<Node> list;
  size_t used = 0;
  for (uint32_t c = 0; c < nch; ++c) list.push_back(Node{w, h, 0, 0, 0xFFFFFFFFu, (int) c});
  std::vector<SqStep> fwd;
  for (const SqueezeParam& p : params) {
    if (p.begin_c + p.num_c > list.size() || p.num_c == 0) return false;
    const uint32_t end = p.begin_c + p.num_c - 1;
    uint32_t offset = p.in_place ? end + 1 : (uint32_t) list.size();
    for (uint32_t c = p.begin_c; c <= end; ++c) {
      Node x = list[c];
      Node a = x, r = x;
      if (p.horizontal) {
        a.w = (x.w + 1) / 2;
        r.w = x.w - a.w;
        a.hs = r.hs = x.hs + 1;
      } else {
        a.h = (x.h + 1) / 2;
        r.h = x.h - a.h;
        a.vs = r.vs = x.vs + 1;
      }
      a.final_channel = r.final_channel = -1;
      a.off = (uint32_t) used;
      used += (size_t) a.w * a.h;
      r.off = (uint32_t) used;
      used += (size_t) r.w * r.h;
      if (used > 0x7FFFFFFFu) return false;
      SqStep s{};
      s.horizontal = p.horizontal ? 1 : 0;
      s.avg_off = a.off;
      s.res_off = r.off;
      s.out_off = x.off;
      s.avg_w = a.w;
      s.avg_h = a.h;
      s.res_w = r.w;
      s.res_h = r.h;
      s.final_channel = x.final_channel >= 0 ? (uint32_t) x.final_channel : 0;
      fwd.push_back(s);
      list[c] = a;
      list.insert(list.begin() + offset + (c - p.begin_c), r);
    }
  }
  out->channels.clear();
  for (const Node& n : list) out->channels.push_back(SqChannel{n.w, n.h, n.hs, n.vs, n.off, 0});
  out->steps.assign(fwd.rbegin(), fwd.rend());
  out->buffer_ints = used;
  return true;
}

}  // namespace jxlb
