// Rescale plan of decodeSampled: what the reference does with RescaleImage (SizeScaler.cpp:38-144) ->
// weave_scale_u8 (/root/reference/weaver/src/scale.rs:100-361) -> pic-scale 0.7.6's fixed-point separable resampler.
//
// The dimension logic (auto sizes -1 / -2, ScaleToFit / ScaleToFill + centre crop, JustResize) is restated from
// weaver/src/scale.rs.  pic-scale itself is a crates.io dependency that is NOT in /root/reference (Cargo.lock:379-382
// pins 0.7.6; only the prebuilt libweaver.a is shipped), so its arithmetic is restated from its behaviour and pinned
// against that binary through the oracle (tests/test_resize_host.py):
//   * scale = in / out (f32); the kernel is stretched by cutoff = max(scale, 1); window size = round(min_kernel * cutoff)
//     taps starting at floor(centre_x - size / 2), clipped to the image; tap weight = kernel(|k - (centre_x - 0.5)| / cutoff)
//     in f32, divided by the (exactly accumulated) sum, quantised to Q15 BY TRUNCATION and SATURATED to int16 (an edge
//     tap of an upscale can exceed 1.0: it becomes 32767 and the row then sums to less than one -- reproduced);
//   * vertical pass first, then horizontal, each: (sum of u8 * w + 2^14) >> 15 saturated to u8 (i32 accumulator); a pass
//     whose size does not change is skipped; equal sizes on both axes return the source untouched;
//   * kernels: Bilinear (2 taps); the BC-spline family, 4 taps -- Cubic and BSpline (B=1, C=0), MitchellNetravali
//     (1/3, 1/3), CatmullRom (0, 1/2), Hermite (0, 0); Nearest is not a convolution: source index =
//     (i * s + s / 2) >> 32 with s = floor(in * 2^32 / out);
//   * sources with alpha (hasAlphaInOrigin): colour is premultiplied with a rounded division by 255 before the passes and
//     divided back afterwards ((p * 255 + a / 2) / a, saturated; a = 0 gives 0) -- except for Nearest and equal sizes;
//   * when ONLY the horizontal pass runs (equal heights), pic-scale 0.7.6 handles the rows four at a time and never
//     processes the remaining height % 4 rows when height >= 4: they stay zero, alpha included (reproduced);
//   * ScaleToFill crops the centre window; pic-scale 0.7.6's crop_with_copy leaves the LAST ROW ZERO when the crop starts
//     at a column > 0 (reproduced: the reference hands that picture to the caller).
// Parity reached: Bilinear and Nearest bit-exact on every size tried; the spline family bit-exact on exact ratios (all of
// BASELINE configs[3]: 7680x4320 -> 1920x1080) and otherwise within one Q15 unit on ~1 tap weight in 10^3 (the f32
// evaluation order of pic-scale's spline polynomial is unknown), i.e. |diff| = 1 on < 0.1 % of the output samples.
//   * Lanczos3 (HANN is mapped to it, SizeScaler.cpp:86-89): sinc(x) sinc(x / 3), 6 taps -- bit-exact against the binary;
//     Bicubic: pic-scale 0.7.6 gives CatmullRom's output bit for bit.
// Refused with kResizeUnsupported, never approximated: Lanczos3 / HANN on sources WITH ALPHA (the filter's strong negative
// lobes push resampled colours above their alpha, and what pic-scale's divide-back does then -- some samples saturate,
// some wrap -- is not pinned), and 16-bit sources (weave_scale_u16 works in f32).
#pragma once
#include <cstdint>
#include <vector>

#include "hd.h"

namespace jxlb {

enum { kResizeOk = 0, kResizeUnsupported = 1, kResizeBadArg = 2 };

struct ResizeAxis {
  uint32_t in_size = 0, out_size = 0, taps = 0;   // taps = window capacity (weights are stored [out_size][taps])
  std::vector<uint32_t> start;                    // first source index per output (Nearest: THE source index)
  std::vector<uint32_t> count;                    // valid taps per output (<= taps)
  std::vector<int16_t> weights;                   // Q15
};

struct ResizePlan {
  uint32_t src_w = 0, src_h = 0;
  uint32_t scaled_w = 0, scaled_h = 0;            // size of the resampled image
  uint32_t crop_x = 0, crop_y = 0, out_w = 0, out_h = 0;  // centre crop applied afterwards (ScaleToFill)
  ResizeAxis v, h;                                // vertical: src_h -> scaled_h; horizontal: src_w -> scaled_w
  bool identity_v = false, identity_h = false;    // pic-scale skips a pass whose size does not change
  bool nearest = false;                           // v.start / h.start are index maps, one gather pass
  bool premultiply = false;                       // premultiply before / divide after the passes
  bool zero_last_row = false;                     // crop quirk (see above)
  uint32_t zero_tail_rows = 0;                    // horizontal-only quirk: the last src_h % 4 rows of the scaled picture stay zero
};

// filter: jxlb_resize_filter (1..10); scale_mode: jxlb_scale_mode (1 Fit, 2 Fill, 3 Resize); has_alpha: hasAlphaInOrigin.
int MakeResizePlan(uint32_t src_w, uint32_t src_h, int32_t req_w, int32_t req_h, int32_t scale_mode, int32_t filter, bool has_alpha,
                   ResizePlan* plan);

// Colour premultiplication around the passes (host + device).
JXLB_HD uint32_t ResizePremul(uint32_t c, uint32_t a) {
  const uint32_t v = c * a + 128;
  return (v + (v >> 8)) >> 8;
}
JXLB_HD uint32_t ResizeUnpremul(uint32_t p, uint32_t a) {
  if (a == 0) return 0;
  const uint32_t v = (p * 255 + a / 2) / a;
  return v > 255 ? 255 : v;
}

// CPU restatement of the passes on an RGBA8 image (tests only; the product runs kernels_resize.cu).
void ResizeRgba8Host(const ResizePlan& plan, const uint8_t* src, uint32_t src_stride, std::vector<uint8_t>* out /* out_w*out_h*4 */);

}  // namespace jxlb
