#include "color_matrix.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

namespace jxlb {

namespace {

struct M3 {
  float v[3][3];
};

// column j = XYZ of primary j with Y = 1 (XyToXYZ, ColorSpaceProfile.h:111-117; the z term is evaluated in double there)
M3 PrimariesXYZ(const float p[3][2]) {
  M3 m;
  for (int j = 0; j < 3; ++j) {
    const float x = p[j][0], y = p[j][1];
    m.v[0][j] = x / y;
    m.v[1][j] = 1.0f;
    m.v[2][j] = (float) ((1.0 - x - y) / y);
  }
  return m;
}

// Eigen's fixed-size 3x3 inverse: cofactors times the reciprocal of the determinant expanded along the first column.
M3 Inverse(const M3& a) {
  auto cof = [&](int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return a.v[i1][j1] * a.v[i2][j2] - a.v[i1][j2] * a.v[i2][j1];
  };
  const float c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
  const float det = (c00 * a.v[0][0] + c10 * a.v[1][0]) + c20 * a.v[2][0];
  const float invdet = 1.0f / det;
  M3 r;
  r.v[0][0] = c00 * invdet;
  r.v[0][1] = c10 * invdet;
  r.v[0][2] = c20 * invdet;
  for (int i = 1; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.v[i][j] = cof(j, i) * invdet;
  return r;
}

M3 Mul(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.v[i][j] = (a.v[i][0] * b.v[0][j] + a.v[i][1] * b.v[1][j]) + a.v[i][2] * b.v[2][j];
  return r;
}

// GamutRgbToXYZ (ColorSpaceProfile.h:131-143)
M3 GamutRgbToXYZ(const float p[3][2], const float w[2]) {
  const M3 xyz = PrimariesXYZ(p);
  const float wx = w[0] / w[1], wy = 1.0f, wz = (float) ((1.0 - w[0] - w[1]) / w[1]);
  const M3 inv = Inverse(xyz);
  float s[3];
  for (int i = 0; i < 3; ++i) s[i] = (inv.v[i][0] * wx + inv.v[i][1] * wy) + inv.v[i][2] * wz;
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.v[i][j] = xyz.v[i][j] * s[j];
  return r;
}

// Trc.cpp: the libavif transfer functions the pass can select
float ToLinear(float g, int tf) {
  switch (tf) {
    case 13:  // sRGB (Trc.cpp:169-179)
      if (g < 0.0f) return 0.0f;
      if (g < 12.92f * 0.0030412825601275209f) return g / 12.92f;
      if (g < 1.0f) return powf((g + 0.0550107189475866f) / 1.0550107189475866f, 2.4f);
      return 1.0f;
    case 1:  // 709 (Trc.cpp:33-43)
      if (g < 0.0f) return 0.0f;
      if (g < 4.5f * 0.018053968510807f) return g / 4.5f;
      if (g < 1.0f) return powf((g + 0.09929682680944f) / 1.09929682680944f, 1.0f / 0.45f);
      return 1.0f;
    case 17:  // DCI -> SMPTE 428 (Trc.cpp:223-225)
      return powf(std::max(g, 0.0f), 2.6f) / 0.91655527974030934f;
    case 16:  // PQ -> extended SDR, 203 nits = 1.0 (Trc.cpp:197-208)
      if (g > 0.0f) {
        const float pg = powf(g, 1.0f / 78.84375f);
        const float num = std::max(pg - 0.8359375f, 0.0f);
        const float den = std::max(18.8515625f - 18.6875f * pg, FLT_MIN);
        const float linear = powf(num / den, 1.0f / 0.1593017578125f);
        return linear * 10000.0f / 203.0f;
      }
      return 0.0f;
    case 18: {  // HLG inverse OETF + OOTF, peak 1000 nits (Trc.cpp:229-246)
      if (g < 0.0f) return 0.0f;
      float linear;
      if (g <= 0.5f) linear = powf((g * g) * (1.0f / 3.0f), 1.2f);
      else linear = powf((expf((g - 0.55991073f) / 0.17883277f) + 0.28466892f) / 12.0f, 1.2f);
      return linear * 1000.0f / 203.0f;
    }
    default:  // gamma -> "Gamma2p2" whatever the coded gamma (JniDecoding.cpp:157-160; Trc.cpp:57-59)
      return powf(std::min(std::max(g, 0.0f), 1.0f), 2.2f);
  }
}

float ToGammaSrgb(float l) {  // Trc.cpp:181-191
  if (l < 0.0f) return 0.0f;
  if (l < 0.0030412825601275209f) return l * 12.92f;
  if (l < 1.0f) return 1.0550107189475866f * powf(l, 1.0f / 2.4f) - 0.0550107189475866f;
  return 1.0f;
}

}  // namespace

bool MakeColorMatrixPlan(const ImageMetadata& md, bool* needed, ColorMatrixPlan* plan, ColorMatrixTables16* tables16) {
  const ColorEncoding& c = md.color;
  *needed = false;
  if (c.want_icc || c.color_space != 0) return true;  // ICC (preferEncoding false) or not RGB: the pass is skipped
  const int tf = c.have_gamma ? 0xFFFF : (int) c.transfer;
  if (!(tf == 16 || tf == 18 || tf == 17 || tf == 1 || tf == 0xFFFF || tf == 13)) return true;  // e.g. linear: skipped
  *needed = true;
  // Rec2408ToneMapper(contentBrightness = intensity target, displayMaxBrightness 250, whitePoint 203) for PQ and HLG
  plan->tonemap = (tf == 16 || tf == 18) ? 1u : 0u;
  {
    const float content = md.intensity_target, display = 250.0f, white = 203.0f;
    const float ld = content / white;
    plan->weight_a = (display / white) / (ld * ld);
    plan->weight_b = 1.0f / (display / white);
    plan->pad = 0;
  }
  float prim[3][2], white[2] = {0.3127f, 0.3290f};
  const float srgb[3][2] = {{0.640f, 0.330f}, {0.300f, 0.600f}, {0.150f, 0.060f}};
  const float bt2020[3][2] = {{0.708f, 0.292f}, {0.170f, 0.797f}, {0.131f, 0.046f}};
  const float p3[3][2] = {{0.68f, 0.32f}, {0.265f, 0.69f}, {0.15f, 0.06f}};
  const float(*src)[2] = nullptr;
  if (c.primaries == 9) src = bt2020;
  else if (c.primaries == 11) src = p3;
  else if (c.primaries == 1) src = srgb;
  if (src) {
    for (int i = 0; i < 3; ++i) prim[i][0] = src[i][0], prim[i][1] = src[i][1];
  } else {
    // custom primaries / white point as libjxl reports them (doubles = coded integer / 1e6), cast to float
    for (int i = 0; i < 3; ++i) {
      prim[i][0] = (float) ((double) c.prim_xy[i][0] * 1e-6);
      prim[i][1] = (float) ((double) c.prim_xy[i][1] * 1e-6);
    }
    if (c.white_point == 1) {
      white[0] = (float) 0.3127;
      white[1] = (float) 0.3290;
    } else if (c.white_point == 2) {
      white[0] = (float) ((double) c.white_xy[0] * 1e-6);
      white[1] = (float) ((double) c.white_xy[1] * 1e-6);
    } else if (c.white_point == 10) {
      white[0] = white[1] = (float) (1.0 / 3.0);
    } else {
      white[0] = (float) 0.314;
      white[1] = (float) 0.351;
    }
  }
  const M3 srcm = GamutRgbToXYZ(prim, white);
  const float d65[2] = {0.3127f, 0.3290f};
  const M3 dstm = GamutRgbToXYZ(srgb, d65);
  const M3 conv = Mul(Inverse(dstm), srcm);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) plan->m[i * 3 + j] = conv.v[i][j];
  for (int j = 0; j < 256; ++j) plan->linearize[j] = ToLinear((float) j * (1.f / 255.f), tf);
  for (int j = 0; j < 2049; ++j)
    plan->gamma[j] = (uint8_t) std::min(std::max(roundf(ToGammaSrgb((float) j * (1.f / 2048.f)) * 255.f), 0.f), 255.f);
  if (tables16) {  // applyColorMatrix16Bit, bitDepth 16 (ColorMatrix.cpp:143-158)
    const float cut = 65535.0f, scale = 1.f / cut;
    for (uint32_t j = 0; j < 65536; ++j) {
      tables16->linearize[j] = ToLinear((float) j * scale, tf);
      tables16->gamma[j] = (uint16_t) std::min(std::max(roundf(ToGammaSrgb((float) j * scale) * cut), 0.f), cut);
    }
  }
  return true;
}

}  // namespace jxlb
