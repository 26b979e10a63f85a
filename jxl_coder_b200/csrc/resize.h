// Rescale plan of decodeSampled: what the reference does with RescaleImage (SizeScaler.cpp:38-144) ->
// weave_scale_u8 (/root/reference/weaver/src/scale.rs:100-361) -> pic-scale 0.7.6's fixed-point separable resampler.
//
// The dimension logic (auto sizes -1 / -2, ScaleToFit / ScaleToFill + centre crop, JustResize) is restated from
// weaver/src/scale.rs.  pic-scale itself is a crates.io dependency that is NOT in /root/reference (Cargo.lock:379-382
// pins 0.7.6; only the prebuilt libweaver.a is shipped), so its arithmetic is restated from its behaviour, pinned
// against that binary through the oracle (tests/test_resize_host.py: bit-exact on every case):
//   * scale = in / out (f32); the kernel is stretched by cutoff = max(scale, 1); window size = round(min_kernel * cutoff)
//     taps starting at floor(centre_x - size / 2), clipped to the image; tap weight = kernel(|k - (centre_x - 0.5)| / cutoff),
//     normalised by the reciprocal of the f32 sum, quantised to Q15 BY TRUNCATION;
//   * vertical pass first, then horizontal, each: (sum of u8 * w + 2^14) >> 15 saturated to u8 (i32 accumulator).
// Supported here: 8-bit sources without alpha, scale >= 1 on both axes, no crop (Fit, Resize, Fill at the source aspect),
// filters Bilinear, MitchellNetravali, CatmullRom and Hermite; everything else reports kResizeUnsupported (never an
// approximate picture).
#pragma once
#include <cstdint>
#include <vector>

namespace jxlb {

enum { kResizeOk = 0, kResizeUnsupported = 1, kResizeBadArg = 2 };

struct ResizeAxis {
  uint32_t in_size = 0, out_size = 0, taps = 0;   // taps = window capacity (weights are stored [out_size][taps])
  std::vector<uint32_t> start;                    // first source index per output
  std::vector<uint32_t> count;                    // valid taps per output (<= taps)
  std::vector<int16_t> weights;                   // Q15
};

struct ResizePlan {
  uint32_t src_w = 0, src_h = 0;
  uint32_t scaled_w = 0, scaled_h = 0;            // size of the resampled image
  uint32_t crop_x = 0, crop_y = 0, out_w = 0, out_h = 0;  // centre crop applied afterwards (ScaleToFill)
  ResizeAxis v, h;                                // vertical: src_h -> scaled_h; horizontal: src_w -> scaled_w
  bool identity_v = false, identity_h = false;    // pic-scale skips a pass whose size does not change
};

// filter: jxlb_resize_filter (1..10); scale_mode: jxlb_scale_mode (1 Fit, 2 Fill, 3 Resize).
int MakeResizePlan(uint32_t src_w, uint32_t src_h, int32_t req_w, int32_t req_h, int32_t scale_mode, int32_t filter, ResizePlan* plan);

// CPU restatement of the two passes on an RGBA8 image (tests only; the product runs kernels_resize.cu).
void ResizeRgba8Host(const ResizePlan& plan, const uint8_t* src, uint32_t src_stride, std::vector<uint8_t>* out /* out_w*out_h*4 */);

}  // namespace jxlb
