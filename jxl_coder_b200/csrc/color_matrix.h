// The colour pass the reference applies below Android 14 (api_level < 34): applyColorMatrix
// (/root/reference/jxlcoder/src/main/cpp/colorspaces/ColorMatrix.cpp:35-119) as driven by decodeSampledImageImpl
// (JniDecoding.cpp:138-228) and getFrameImpl (JxlAnimatedDecoderCoordinator.cpp:182-265): for an RGB image whose colour
// encoding is an enum (not ICC) with transfer function sRGB / 709 / gamma / DCI, every 8-bit pixel goes through
//   256-entry linearisation LUT (toLinear, colorspaces/Trc.cpp:265-296)  ->  3x3 matrix source primaries -> Rec.709
//   (GamutRgbToXYZ, colorspaces/ColorSpaceProfile.h:131-143; dst^-1 * src)  ->  clamp [0,1], * 2048, truncate  ->
//   2049-entry sRGB LUT (toGamma, round(v * 255)); alpha untouched.
// It runs even for plain sRGB images (identity-like matrix: a lossy 2048-level requantisation).
// The tables and the matrix are built on the host in f32 with the reference's operation order (Eigen's 3x3 cofactor
// inverse); kernels_post.cu applies them.  PQ / HLG sources (Rec.2408 tone mapping) and 16-bit sources
// (applyColorMatrix16Bit) are not restated: MakeColorMatrixPlan returns false and the decoder reports JXLB_UNSUPPORTED.
#pragma once
#include <cstdint>

#include "frame_parser.h"

namespace jxlb {

struct ColorMatrixPlan {
  float linearize[256];
  uint8_t gamma[2049 + 3];
  float m[9];
};

// needed: the reference's condition for running the pass (enum encoding, RGB, one of the listed transfer functions).
// Returns false when the pass is needed but not covered here.
bool MakeColorMatrixPlan(const ImageMetadata& md, bool* needed, ColorMatrixPlan* plan);

// CPU restatement of the pass on RGBA8 rows (tests only).
void ApplyColorMatrixHost(const ColorMatrixPlan& plan, uint8_t* rgba, uint32_t stride, uint32_t width, uint32_t height);

}  // namespace jxlb
