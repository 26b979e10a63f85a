// Alpha association + pixel-format packing, host + device.  Restates, bit-exactly (integer arithmetic; the only float
// step is u8|u16 -> half with round-to-nearest-even), the reference's post-decode reformat:
//   ReformatColorConfig            /root/reference/jxlcoder/src/main/cpp/ReformatBitmap.cpp:46-193
//   coder::AssociateAlphaRgba8/16  imagebit/RGBAlpha.cpp:67-117        (c * a / 255, truncating; 16-bit: / (2^depth - 1))
//   coder::Rgba8ToF16              imagebit/Rgba8ToF16.cpp:44-138      (half(c * (1/255)), optional second premultiply)
//   coder::RgbaU16ToF              imagebit/RgbaU16toHF.cpp:42-144     (half(c * (1/65535)))
//   coder::Rgba8ToRGBA1010102 / Rgba16ToRGBA1010102   imagebit/Rgb1010102.cpp:177-248   (r << 2, a >> 6; >> (depth-10))
//   coder::Rgba8To565 / Rgba16To565                   imagebit/Rgb565.cpp:99-160
//   coder::Rgba16ToRgba8           imagebit/Rgba16.cpp:32-68           (>> (depth - 8))
// including the reference's double premultiplication for 8-bit sources converted to F16 / 565 / 1010102
// (SURVEY.md App. D item 2).
#pragma once
#include "hd.h"
#ifdef __CUDACC__
#include <cuda_fp16.h>
#endif

namespace jxlb {

struct PackParams {
  const uint8_t* src;      // RGBA interleaved u8 or u16 (decode stage output)
  uint32_t src_stride;     // bytes
  uint8_t* dst;
  uint32_t dst_stride;     // bytes
  uint32_t width, height;
  uint32_t src16;          // source samples are uint16
  uint32_t depth;          // 8 or 16 (bit depth the reference tracks)
  uint32_t format;         // jxlb_format: 0 8888, 1 F16, 2 565, 3 1010102; 4 (internal): the samples as they are, RGBA8 or
                           // RGBA16 -- the staging image in front of the orientation / rescale / colour passes
  uint32_t associate;      // step 1: !alphaPremultiplied && hasAlphaInOrigin
  uint32_t attenuate;      // step 2 (8-bit sources only): !alphaPremultiplied
};

JXLB_HD uint16_t FloatToHalfBits(float f) {
#ifdef __CUDA_ARCH__
  return __half_as_ushort(__float2half_rn(f));
#else
  // round-to-nearest-even float -> half for the non-negative finite values that occur here
  union { float f; uint32_t u; } v;
  v.f = f;
  uint32_t x = v.u;
  uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7FFFFFFFu;
  if (x >= 0x47800000u) return (uint16_t) (sign | 0x7C00u);            // overflow -> inf
  if (x < 0x38800000u) {                                               // subnormal half
    if (x < 0x33000000u) return (uint16_t) sign;
    uint32_t shift = 113 - (x >> 23);
    uint32_t mant = (x & 0x7FFFFFu) | 0x800000u;
    uint32_t half = mant >> (shift + 13);
    uint32_t rem = mant & ((1u << (shift + 13)) - 1);
    uint32_t halfway = 1u << (shift + 12);
    if (rem > halfway || (rem == halfway && (half & 1))) ++half;
    return (uint16_t) (sign | half);
  }
  uint32_t half = ((x - 0x38000000u) >> 13);
  uint32_t rem = x & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) ++half;
  return (uint16_t) (sign | half);
#endif
}

// Packs one decoded pixel (r, g, b, a at the decode stage's depth: 8-bit, or 16-bit when p.src16) into the output.
JXLB_HD void PackRgba(const PackParams& p, uint32_t x, uint32_t y, uint32_t r, uint32_t g, uint32_t b, uint32_t a) {
  if (p.format == 4) {  // staging: no association, no conversion
    uint8_t* drow = p.dst + (size_t) y * p.dst_stride;
    if (p.src16) {
      uint16_t* o = reinterpret_cast<uint16_t*>(drow) + 4 * (size_t) x;
      o[0] = (uint16_t) r;
      o[1] = (uint16_t) g;
      o[2] = (uint16_t) b;
      o[3] = (uint16_t) a;
    } else {
      reinterpret_cast<uint32_t*>(drow)[x] = (r & 0xFFu) | ((g & 0xFFu) << 8) | ((b & 0xFFu) << 16) | ((a & 0xFFu) << 24);
    }
    return;
  }
  if (p.src16) {
    if (p.associate) {
      const uint32_t maxc = (1u << p.depth) - 1;
      r = r * a / maxc;
      g = g * a / maxc;
      b = b * a / maxc;
    }
    uint8_t* drow = p.dst + (size_t) y * p.dst_stride;
    switch (p.format) {
      case 0: {
        const uint32_t d = p.depth - 8;
        uint8_t* o = drow + 4 * (size_t) x;
        o[0] = (uint8_t) (r >> d);
        o[1] = (uint8_t) (g >> d);
        o[2] = (uint8_t) (b >> d);
        o[3] = (uint8_t) (a >> d);
        break;
      }
      case 1: {
        const float scale = 1.0f / (float) ((1u << p.depth) - 1);
        uint16_t* o = reinterpret_cast<uint16_t*>(drow) + 4 * (size_t) x;
        o[0] = FloatToHalfBits((float) r * scale);
        o[1] = FloatToHalfBits((float) g * scale);
        o[2] = FloatToHalfBits((float) b * scale);
        o[3] = FloatToHalfBits((float) a * scale);
        break;
      }
      case 2: {
        const uint32_t gd = p.depth - 8 + 2, rbd = p.depth - 8 + 3;
        reinterpret_cast<uint16_t*>(drow)[x] = (uint16_t) (((r >> rbd) << 11) | ((g >> gd) << 5) | (b >> rbd));
        break;
      }
      default: {
        const uint32_t d = p.depth - 10, ad = p.depth - 2;
        reinterpret_cast<uint32_t*>(drow)[x] = ((a >> ad) & 3u) << 30 | ((b >> d) & 0x3FFu) << 20 | ((g >> d) & 0x3FFu) << 10 | ((r >> d) & 0x3FFu);
        break;
      }
    }
    return;
  }
  if (p.associate) {
    r = r * a / 255u;
    g = g * a / 255u;
    b = b * a / 255u;
  }
  uint8_t* drow = p.dst + (size_t) y * p.dst_stride;
  if (p.format == 0) {
    reinterpret_cast<uint32_t*>(drow)[x] = (r & 0xFFu) | ((g & 0xFFu) << 8) | ((b & 0xFFu) << 16) | ((a & 0xFFu) << 24);  // R,G,B,A bytes
    return;
  }
  if (p.attenuate) {  // the converters premultiply again (reference quirk)
    r = r * a / 255u;
    g = g * a / 255u;
    b = b * a / 255u;
  }
  if (p.format == 1) {
    const float scale = 1.0f / 255.0f;
    uint16_t* o = reinterpret_cast<uint16_t*>(drow) + 4 * (size_t) x;
    o[0] = FloatToHalfBits((float) r * scale);
    o[1] = FloatToHalfBits((float) g * scale);
    o[2] = FloatToHalfBits((float) b * scale);
    o[3] = FloatToHalfBits((float) a * scale);
  } else if (p.format == 2) {
    reinterpret_cast<uint16_t*>(drow)[x] = (uint16_t) (((r >> 3) << 11) | ((g >> 2) << 5) | (b >> 3));
  } else {
    reinterpret_cast<uint32_t*>(drow)[x] = ((a >> 6) << 30) | ((b << 2) << 20) | ((g << 2) << 10) | (r << 2);
  }
}

JXLB_HD void PackPixel(const PackParams& p, uint32_t x, uint32_t y) {
  if (p.src16) {
    const uint16_t* s = reinterpret_cast<const uint16_t*>(p.src + (size_t) y * p.src_stride) + 4 * (size_t) x;
    PackRgba(p, x, y, s[0], s[1], s[2], s[3]);
  } else {
    const uint8_t* s = p.src + (size_t) y * p.src_stride + 4 * (size_t) x;
    PackRgba(p, x, y, s[0], s[1], s[2], s[3]);
  }
}

JXLB_HD uint32_t FormatBytesPerPixel(uint32_t format) { return format == 1 ? 8u : format == 2 ? 2u : 4u; }  // not for the staging format

}  // namespace jxlb
