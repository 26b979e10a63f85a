// Host-side setup of the XYB -> output colour transform (SURVEY.md App. B.7 "XYB -> RGB"): inverse opsin matrix scaled
// by 255 / intensity_target, linear-sRGB -> target-primaries matrix from the xy chromaticities, output transfer function.
#include "color_params.h"

#include <cmath>

namespace jxlb {

namespace {

void RgbToXyz(const double prim[3][2], const double wp[2], double m[9]) {
  // columns = primaries' XYZ scaled so that R = G = B = 1 maps to the white point
  double P[9] = {prim[0][0] / prim[0][1], prim[1][0] / prim[1][1], prim[2][0] / prim[2][1],
                 1, 1, 1,
                 (1 - prim[0][0] - prim[0][1]) / prim[0][1], (1 - prim[1][0] - prim[1][1]) / prim[1][1], (1 - prim[2][0] - prim[2][1]) / prim[2][1]};
  double W[3] = {wp[0] / wp[1], 1, (1 - wp[0] - wp[1]) / wp[1]};
  // solve P * S = W
  double det = P[0] * (P[4] * P[8] - P[5] * P[7]) - P[1] * (P[3] * P[8] - P[5] * P[6]) + P[2] * (P[3] * P[7] - P[4] * P[6]);
  double inv[9] = {(P[4] * P[8] - P[5] * P[7]) / det, (P[2] * P[7] - P[1] * P[8]) / det, (P[1] * P[5] - P[2] * P[4]) / det,
                   (P[5] * P[6] - P[3] * P[8]) / det, (P[0] * P[8] - P[2] * P[6]) / det, (P[2] * P[3] - P[0] * P[5]) / det,
                   (P[3] * P[7] - P[4] * P[6]) / det, (P[1] * P[6] - P[0] * P[7]) / det, (P[0] * P[4] - P[1] * P[3]) / det};
  double S[3];
  for (int i = 0; i < 3; ++i) S[i] = inv[3 * i] * W[0] + inv[3 * i + 1] * W[1] + inv[3 * i + 2] * W[2];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) m[3 * r + c] = P[3 * r + c] * S[c];
}

void Invert3(const double a[9], double inv[9]) {
  double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
  double r[9] = {(a[4] * a[8] - a[5] * a[7]) / det, (a[2] * a[7] - a[1] * a[8]) / det, (a[1] * a[5] - a[2] * a[4]) / det,
                 (a[5] * a[6] - a[3] * a[8]) / det, (a[0] * a[8] - a[2] * a[6]) / det, (a[2] * a[3] - a[0] * a[5]) / det,
                 (a[3] * a[7] - a[4] * a[6]) / det, (a[1] * a[6] - a[0] * a[7]) / det, (a[0] * a[4] - a[1] * a[3]) / det};
  for (int i = 0; i < 9; ++i) inv[i] = r[i];
}

}  // namespace

int MakeColorParams(const ImageMetadata& md, ColorParams* cp, std::string* err) {
  static const double kOpsinInv[9] = {11.031566901960783, -9.866943921568629, -0.16462299647058826,
                                      -3.254147380392157, 4.418770392156863, -0.16462299647058826,
                                      -3.6588512862745097, 2.7129230470588235, 1.9459282392156863};
  const float it = md.intensity_target > 0 ? md.intensity_target : 255.0f;
  double fold[9];  // (sRGB -> target primaries) x inverse opsin matrix
  for (int i = 0; i < 9; ++i) {
    fold[i] = kOpsinInv[i];
    cp->to_target[i] = (i % 4 == 0) ? 1.0f : 0.0f;
  }
  cp->apply_primaries = 0;
  cp->transfer = md.color.have_gamma ? 0xFFFFu : md.color.transfer;
  cp->gamma = md.color.have_gamma ? (float) (md.color.gamma_u24 * 1e-7) : 1.0f;
  cp->pq_scale = it * (1.0f / 10000.0f);  // TF_PQ's display_scaling_factor_to_10000_nits_, in float like libjxl
  cp->grey = md.color.color_space == 1;
  const ColorEncoding& ce = md.color;
  if (ce.color_space != 0 && ce.color_space != 1) {
    if (err) *err = "unsupported colour space";
    return kParseUnsupported;
  }
  switch (cp->transfer) {
    case 13: case 8: case 1: case 16: case 17: case 0xFFFF: break;
    default:
      if (err) *err = "unsupported transfer function";
      return kParseUnsupported;
  }
  if (ce.white_point != 1) {
    if (err) *err = "non-D65 white point";
    return kParseUnsupported;
  }
  if (ce.color_space == 0 && ce.primaries != 1) {
    double prim[3][2];
    if (ce.primaries == 9) {
      const double p[3][2] = {{0.708, 0.292}, {0.170, 0.797}, {0.131, 0.046}};
      for (int i = 0; i < 3; ++i) prim[i][0] = p[i][0], prim[i][1] = p[i][1];
    } else if (ce.primaries == 11) {
      const double p[3][2] = {{0.680, 0.320}, {0.265, 0.690}, {0.150, 0.060}};
      for (int i = 0; i < 3; ++i) prim[i][0] = p[i][0], prim[i][1] = p[i][1];
    } else if (ce.primaries == 2) {
      for (int i = 0; i < 3; ++i) prim[i][0] = ce.prim_xy[i][0] * 1e-6, prim[i][1] = ce.prim_xy[i][1] * 1e-6;
    } else {
      if (err) *err = "unsupported primaries";
      return kParseUnsupported;
    }
    // libjxl's sRGB primaries are the ones of the sRGB ICC profile, not the rounded ITU values (3e-5 apart: enough to move
    // near-black PQ samples by several codes); found by fitting the reference's 16-bit Rec.2100 output against the XYB
    // planes (agreement 3e-6 = the fit's noise floor, against 7e-5 with 0.64 / 0.33 ...)
    const double srgb[3][2] = {{0.639998686, 0.330010138}, {0.300003784, 0.600003357}, {0.150002046, 0.059997204}};
    const double d65[2] = {0.3127, 0.3290};
    double ms[9], mt[9], mti[9], tt[9];
    RgbToXyz(srgb, d65, ms);
    RgbToXyz(prim, d65, mt);
    Invert3(mt, mti);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double v = 0;
        for (int k = 0; k < 3; ++k) v += mti[3 * r + k] * ms[3 * k + c];
        cp->to_target[3 * r + c] = (float) v;
        tt[3 * r + c] = v;
      }
    cp->apply_primaries = 1;
    // libjxl folds the primaries conversion into the inverse opsin matrix (one 3x3 per pixel instead of two)
    double m[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double v = 0;
        for (int k = 0; k < 3; ++k) v += tt[3 * r + k] * kOpsinInv[3 * k + c];
        m[3 * r + c] = v;
      }
    for (int i = 0; i < 9; ++i) fold[i] = m[i];
  }
  // InitSIMDInverseMatrix: float matrix entry times the float 255 / intensity_target
  for (int i = 0; i < 9; ++i) cp->opsin_inv[i] = (float) fold[i] * (255.0f / it);
  return kParseOk;
}

}  // namespace jxlb
