/* TEST INFRASTRUCTURE (oracle) — plain-C restatement of the reference's post-decode reformat, used to check the CUDA
 * pack kernel where oracle/_ref (the reference's own compiled sources) is unavailable.  Not shipped, not measured.
 *   AssociateAlphaRgba8          /root/reference/jxlcoder/src/main/cpp/imagebit/RGBAlpha.cpp:67-90
 *   Rgba8ToRGBA1010102           imagebit/Rgb1010102.cpp:177-211
 *   Rgba8To565                   imagebit/Rgb565.cpp:99-130
 *   Rgba16ToRgba8                imagebit/Rgba16.cpp:32-68
 * parity: pinned against oracle/_ref in tests/test_postproc_port.py. */
#include <stdint.h>
#include <stddef.h>

void port_associate_alpha_rgba8(const uint8_t* src, uint8_t* dst, size_t npix) {
  for (size_t i = 0; i < npix; ++i) {
    uint16_t a = src[4 * i + 3];
    dst[4 * i + 0] = (uint8_t) ((uint16_t) src[4 * i + 0] * a / 255);
    dst[4 * i + 1] = (uint8_t) ((uint16_t) src[4 * i + 1] * a / 255);
    dst[4 * i + 2] = (uint8_t) ((uint16_t) src[4 * i + 2] * a / 255);
    dst[4 * i + 3] = (uint8_t) a;
  }
}

void port_rgba8_to_1010102(const uint8_t* src, uint32_t* dst, size_t npix, int attenuate) {
  for (size_t i = 0; i < npix; ++i) {
    uint32_t r = src[4 * i], g = src[4 * i + 1], b = src[4 * i + 2], a = src[4 * i + 3];
    if (attenuate) {
      r = r * a / 255;
      g = g * a / 255;
      b = b * a / 255;
    }
    dst[i] = ((a >> 6) << 30) | ((b << 2) << 20) | ((g << 2) << 10) | (r << 2);
  }
}

void port_rgba8_to_565(const uint8_t* src, uint16_t* dst, size_t npix, int attenuate) {
  for (size_t i = 0; i < npix; ++i) {
    uint32_t r = src[4 * i], g = src[4 * i + 1], b = src[4 * i + 2], a = src[4 * i + 3];
    if (attenuate) {
      r = r * a / 255;
      g = g * a / 255;
      b = b * a / 255;
    }
    dst[i] = (uint16_t) (((r >> 3) << 11) | ((g >> 2) << 5) | (b >> 3));
  }
}

void port_rgba16_to_rgba8(const uint16_t* src, uint8_t* dst, size_t npix, int depth) {
  int d = depth - 8;
  for (size_t i = 0; i < 4 * npix; ++i) dst[i] = (uint8_t) (src[i] >> d);
}
