// TEST-ONLY: CPU restatement of the api_level < 34 colour pass (color_matrix.h) on host rows.  Compiled into tests/hostemu,
// NOT into libjxlb200.so: the product applies the plan with ColorMatrixKernel (kernels_post.cu) and has no CPU pixel path.
#include <algorithm>
#include <cmath>
#include <vector>

#include "color_matrix.h"

namespace jxlb {

namespace {
// Rec2408ToneMapper::transferTone on one row of linear RGB triplets, with its pointer bug: from the first pixel of zero
// luminance on, nothing is tone-mapped.
void ToneMapRow(const ColorMatrixPlan& p, float* rgb, uint32_t width) {
  for (uint32_t x = 0; x < width; ++x) {
    float* t = rgb + 3 * x;
    const float light = 0.2627f * t[0] + 0.6780f * t[1] + 0.0593f * t[2];
    if (light == 0) return;
    const float scale = (1.f + p.weight_a * light) / (1.f + p.weight_b * light);
    t[0] = std::min(t[0] * scale, 1.f);
    t[1] = std::min(t[1] * scale, 1.f);
    t[2] = std::min(t[2] * scale, 1.f);
  }
}
}  // namespace

void ApplyColorMatrixHost(const ColorMatrixPlan& p, uint8_t* rgba, uint32_t stride, uint32_t width, uint32_t height) {
  std::vector<float> rowv((size_t) width * 3);
  for (uint32_t y = 0; y < height; ++y) {
    uint8_t* row = rgba + (size_t) y * stride;
    for (uint32_t x = 0; x < width; ++x)
      for (int c = 0; c < 3; ++c) rowv[3 * x + c] = p.linearize[row[4 * x + c]];
    if (p.tonemap) ToneMapRow(p, rowv.data(), width);
    for (uint32_t x = 0; x < width; ++x, row += 4) {
      const float r = rowv[3 * x], g = rowv[3 * x + 1], b = rowv[3 * x + 2];
      const float v[3] = {r * p.m[0] + g * p.m[1] + b * p.m[2], r * p.m[3] + g * p.m[4] + b * p.m[5], r * p.m[6] + g * p.m[7] + b * p.m[8]};
      for (int c = 0; c < 3; ++c) {
        const uint32_t idx = std::min<uint32_t>((uint16_t) (std::min(std::max(v[c], 0.f), 1.0f) * 2048.f), 2048u);
        row[c] = p.gamma[idx];
      }
    }
  }
}

void ApplyColorMatrixHost16(const ColorMatrixPlan& p, const ColorMatrixTables16& t, uint16_t* rgba, uint32_t stride_bytes, uint32_t width,
                            uint32_t height) {
  std::vector<float> rowv((size_t) width * 3);
  for (uint32_t y = 0; y < height; ++y) {
    uint16_t* row = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(rgba) + (size_t) y * stride_bytes);
    for (uint32_t x = 0; x < width; ++x)
      for (int c = 0; c < 3; ++c) rowv[3 * x + c] = t.linearize[row[4 * x + c]];
    if (p.tonemap) ToneMapRow(p, rowv.data(), width);
    for (uint32_t x = 0; x < width; ++x, row += 4) {
      const float r = rowv[3 * x], g = rowv[3 * x + 1], b = rowv[3 * x + 2];
      const float v[3] = {r * p.m[0] + g * p.m[1] + b * p.m[2], r * p.m[3] + g * p.m[4] + b * p.m[5], r * p.m[6] + g * p.m[7] + b * p.m[8]};
      for (int c = 0; c < 3; ++c) {
        const uint32_t idx = std::min<uint32_t>((uint16_t) (std::min(std::max(v[c], 0.f), 1.0f) * 65535.f), 65535u);
        row[c] = t.gamma[idx];
      }
    }
  }
}

}  // namespace jxlb
