// TEST-ONLY: CPU restatement of the rescale passes (resize.h) on host rows.  Compiled into tests/hostemu, NOT into
// libjxlb200.so: the product runs the plan with kernels_resize.cu and has no CPU pixel path.
#include "resize.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace jxlb {

void ResizeRgba8Host(const ResizePlan& p, const uint8_t* src, uint32_t src_stride, std::vector<uint8_t>* out) {
  std::vector<uint8_t> scaled((size_t) p.scaled_h * p.scaled_w * 4);
  if (p.nearest) {
    for (uint32_t y = 0; y < p.scaled_h; ++y)
      for (uint32_t x = 0; x < p.scaled_w; ++x)
        memcpy(&scaled[((size_t) y * p.scaled_w + x) * 4], src + (size_t) p.v.start[y] * src_stride + (size_t) p.h.start[x] * 4, 4);
  } else {
    std::vector<uint8_t> pre((size_t) p.src_h * p.src_w * 4);
    for (uint32_t y = 0; y < p.src_h; ++y)
      for (uint32_t x = 0; x < p.src_w; ++x) {
        const uint8_t* s = src + (size_t) y * src_stride + (size_t) x * 4;
        uint8_t* d = &pre[((size_t) y * p.src_w + x) * 4];
        for (int c = 0; c < 3; ++c) d[c] = p.premultiply ? (uint8_t) ResizePremul(s[c], s[3]) : s[c];
        d[3] = s[3];
      }
    std::vector<uint8_t> mid((size_t) p.scaled_h * p.src_w * 4);
    for (uint32_t y = 0; y < p.scaled_h; ++y)
      for (uint32_t x = 0; x < p.src_w * 4; ++x) {
        if (p.identity_v) {
          mid[(size_t) y * p.src_w * 4 + x] = pre[(size_t) y * p.src_w * 4 + x];
          continue;
        }
        int32_t acc = 1 << 14;
        const int16_t* w = &p.v.weights[(size_t) y * p.v.taps];
        for (uint32_t t = 0; t < p.v.count[y]; ++t) acc += (int32_t) w[t] * pre[(size_t) (p.v.start[y] + t) * p.src_w * 4 + x];
        acc >>= 15;
        mid[(size_t) y * p.src_w * 4 + x] = (uint8_t) std::min(std::max(acc, 0), 255);
      }
    for (uint32_t y = 0; y < p.scaled_h; ++y)
      for (uint32_t x = 0; x < p.scaled_w; ++x) {
        uint8_t* d = &scaled[((size_t) y * p.scaled_w + x) * 4];
        for (uint32_t c = 0; c < 4; ++c) {
          if (p.identity_h) {
            d[c] = mid[((size_t) y * p.src_w + x) * 4 + c];
            continue;
          }
          int32_t acc = 1 << 14;
          const int16_t* w = &p.h.weights[(size_t) x * p.h.taps];
          for (uint32_t t = 0; t < p.h.count[x]; ++t) acc += (int32_t) w[t] * mid[((size_t) y * p.src_w + p.h.start[x] + t) * 4 + c];
          acc >>= 15;
          d[c] = (uint8_t) std::min(std::max(acc, 0), 255);
        }
        if (p.premultiply)
          for (int c = 0; c < 3; ++c) d[c] = (uint8_t) ResizeUnpremul(d[c], d[3]);
        if (y + p.zero_tail_rows >= p.scaled_h) d[0] = d[1] = d[2] = d[3] = 0;
      }
  }
  out->assign((size_t) p.out_w * p.out_h * 4, 0);
  for (uint32_t y = 0; y < p.out_h; ++y) {
    if (p.zero_last_row && y + 1 == p.out_h) break;
    std::copy_n(&scaled[((size_t) (y + p.crop_y) * p.scaled_w + p.crop_x) * 4], (size_t) p.out_w * 4, &(*out)[(size_t) y * p.out_w * 4]);
  }
}

}  // namespace jxlb
