// The colour pass the reference applies below Android 14 (api_level < 34): applyColorMatrix
// (/root/reference/jxlcoder/src/main/cpp/colorspaces/ColorMatrix.cpp:35-119) as driven by decodeSampledImageImpl
// (JniDecoding.cpp:138-228) and getFrameImpl (JxlAnimatedDecoderCoordinator.cpp:182-265): for an RGB image whose colour
// encoding is an enum (not ICC) with transfer function sRGB / 709 / gamma / DCI, every 8-bit pixel goes through
//   256-entry linearisation LUT (toLinear, colorspaces/Trc.cpp:265-296)  ->  3x3 matrix source primaries -> Rec.709
//   (GamutRgbToXYZ, colorspaces/ColorSpaceProfile.h:131-143; dst^-1 * src)  ->  clamp [0,1], * 2048, truncate  ->
//   2049-entry sRGB LUT (toGamma, round(v * 255)); alpha untouched.
// It runs even for plain sRGB images (identity-like matrix: a lossy 2048-level requantisation).
// PQ and HLG sources are linearised to "extended SDR" (203 nits = 1.0, Trc.cpp:197-252) and tone-mapped in between with
// Rec2408ToneMapper::transferTone (Rec2408ToneMapper.cpp:80-100; content brightness = the image's intensity target,
// display 250 nits, white 203 nits) -- INCLUDING its row bug: a pixel of zero luminance hits `continue` without advancing
// the pointer, so that pixel and everything after it in the row stays un-tone-mapped.
// Sources deeper than 8 bits take applyColorMatrix16Bit (ColorMatrix.cpp:121-219): the same with 2^16-entry tables.
// The tables and the matrix are built on the host in f32 with the reference's operation order (Eigen's 3x3 cofactor
// inverse); kernels_post.cu applies them.
#pragma once
#include <cstdint>

#include "frame_parser.h"

namespace jxlb {

struct ColorMatrixPlan {
  float linearize[256];
  uint8_t gamma[2049 + 3];
  float m[9];
  uint32_t tonemap;          // PQ / HLG: Rec2408ToneMapper between the linearisation and the matrix
  float weight_a, weight_b;  // its two constants
  uint32_t pad;
};
// 16-bit variant (applyColorMatrix16Bit with bitDepth 16): follows the plan in the const region when the source is deeper
// than 8 bits.
struct ColorMatrixTables16 {
  float linearize[65536];
  uint16_t gamma[65536];
};

// needed: the reference's condition for running the pass (enum encoding, RGB, one of the listed transfer functions).
// tables16: filled when non-null (sources deeper than 8 bits).  Returns false when the pass is needed but not covered here.
bool MakeColorMatrixPlan(const ImageMetadata& md, bool* needed, ColorMatrixPlan* plan, ColorMatrixTables16* tables16 = nullptr);

// CPU restatement of the pass on RGBA8 / RGBA16 rows (tests only; the product applies the plan with kernels_post.cu).
void ApplyColorMatrixHost(const ColorMatrixPlan& plan, uint8_t* rgba, uint32_t stride, uint32_t width, uint32_t height);
void ApplyColorMatrixHost16(const ColorMatrixPlan& plan, const ColorMatrixTables16& t, uint16_t* rgba, uint32_t stride_bytes, uint32_t width,
                            uint32_t height);

}  // namespace jxlb
