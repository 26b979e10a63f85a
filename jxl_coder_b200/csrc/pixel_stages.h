// Per-pixel stages after the inverse transforms, host + device: Gaborish, EPF stages, XYB -> RGBA (u8 dithered / u16),
// modular (lossless) samples -> RGBA.  One call = one output pixel; the kernels map one thread per pixel.
// Output matches what the reference asks libjxl for: interleaved RGBA, JxlPixelFormat{4, UINT8|UINT16, native, 0}
// (/root/reference/jxlcoder/src/main/cpp/interop/JxlDecoding.cpp:63,96), alpha not premultiplied.
#pragma once
#include "numeric.h"

namespace jxlb {

struct OutputDesc {
  uint8_t* data;           // RGBA interleaved
  uint32_t stride_bytes;
  uint32_t bits16;         // 0: uint8 samples, 1: uint16 samples
  int32_t alpha_channel;   // index into the frame's modular image, or -1
  uint32_t alpha_bits;     // bit depth of the alpha samples
  uint32_t color_bits;     // bit depth of modular colour samples (modular frames)
  uint32_t alpha_float = 0;  // the alpha plane holds floats in [0, 1] (an upsampled alpha): quantised -- and, at 8 bits, dithered -- like colour
};

JXLB_HD Planes3 ViewPlanes(const FrameDev& f, const float* base) {
  Planes3 v;
  const size_t plane = (size_t) f.plane_h * f.plane_stride;
  v.p[0] = base;
  v.p[1] = base + plane;
  v.p[2] = base + 2 * plane;
  v.w = (int) f.width;
  v.h = (int) f.height;
  v.stride = (int) f.plane_stride;
  return v;
}

JXLB_HD void StageGaborish(const FrameDev& f, const float* src, float* dst, int x, int y) {
  const Planes3 im = ViewPlanes(f, src);
  const size_t plane = (size_t) f.plane_h * f.plane_stride, o = (size_t) y * f.plane_stride + x;
  for (int c = 0; c < 3; ++c) dst[c * plane + o] = GaborishSample(im, c, x, y, f.rf.gab_w1[c], f.rf.gab_w2[c]);
}

JXLB_HD void StageEpf(const FrameDev& f, const NumericTables& nt, int stage, const float* src, float* dst, int x, int y) {
  const Planes3 im = ViewPlanes(f, src);
  const size_t plane = (size_t) f.plane_h * f.plane_stride, o = (size_t) y * f.plane_stride + x;
  const size_t ci = (size_t) (y >> 3) * f.w8 + (x >> 3);
  const float inv_sigma = EpfInvSigma(f, f.cell_hfmul[ci], f.cell_sharp[ci]);
  float out[3];
  EpfPixel(im, f.rf, nt.rcp11, stage, x, y, x, y, inv_sigma, out);
  dst[o] = out[0];
  dst[plane + o] = out[1];
  dst[2 * plane + o] = out[2];
}

JXLB_HD uint32_t ScaleSample(int32_t v, uint32_t bits, uint32_t maxout) {
  const uint32_t maxin = (1u << bits) - 1;
  if (v < 0) v = 0;
  if ((uint32_t) v > maxin) v = (int32_t) maxin;
  if (maxin == maxout) return (uint32_t) v;
  const float fv = (float) v * (1.0f / (float) maxin) * (float) maxout;
  return (uint32_t) rintf(fv);
}

JXLB_HD void StoreRgba(const OutputDesc& out, int x, int y, uint32_t r, uint32_t g, uint32_t b, uint32_t a) {
  if (out.bits16) {
    uint16_t* p = reinterpret_cast<uint16_t*>(out.data + (size_t) y * out.stride_bytes) + 4 * (size_t) x;
    p[0] = (uint16_t) r;
    p[1] = (uint16_t) g;
    p[2] = (uint16_t) b;
    p[3] = (uint16_t) a;
  } else {
    uint8_t* p = out.data + (size_t) y * out.stride_bytes + 4 * (size_t) x;
    p[0] = (uint8_t) r;
    p[1] = (uint8_t) g;
    p[2] = (uint8_t) b;
    p[3] = (uint8_t) a;
  }
}

JXLB_HD uint32_t AlphaAt(const FrameDev& f, const OutputDesc& out, int x, int y) {
  const uint32_t maxout = out.bits16 ? 65535u : 255u;
  if (out.alpha_channel < 0) return maxout;
  const int32_t a = f.mod[(size_t) out.alpha_channel * f.height * f.mod_stride + (size_t) y * f.mod_stride + x];
  return ScaleSample(a, out.alpha_bits, maxout);
}

// libjxl dithers after it has FLIPPED the picture for its orientation but before the transposition of orientations 5..8
// (found by trying all index maps against the reference, tests/test_orientation_host.py): the 32x32 pattern is indexed
// by the position in the flipped, untransposed image.  (x, y) are coded coordinates.
JXLB_HD uint32_t DitherIndex(const FrameDev& f, uint32_t x, uint32_t y) {
  const uint32_t o = f.orientation;
  const bool flip_x = o == 2 || o == 3 || o == 7 || o == 8, flip_y = o == 3 || o == 4 || o == 6 || o == 7;
  const uint32_t ox = flip_x ? f.width - 1 - x : x, oy = flip_y ? f.height - 1 - y : y;
  return ((oy + f.dither_y0) & 31) * 32 + ((ox + f.dither_x0) & 31);
}

JXLB_HD void StageColorToRgba(const FrameDev& f, const ColorParams& cp, const NumericTables& nt, const float* src,
                              const OutputDesc& out, int x, int y) {
  const size_t plane = (size_t) f.plane_h * f.plane_stride, o = (size_t) y * f.plane_stride + x;
  float rgb[3];
  XybToEncodedRgb(src[o], src[plane + o], src[2 * plane + o], cp, rgb);
  uint32_t v[3];
  // an upsampled alpha is a float in [0, 1] that goes through the same output conversion as colour
  const float af = out.alpha_float ? reinterpret_cast<const float*>(f.mod)[(size_t) y * f.mod_stride + x] : 0.0f;
  uint32_t alpha;
  if (out.bits16) {
    for (int c = 0; c < 3; ++c) {
      float s = rgb[c] * 65535.0f;
      s = s < 0.0f ? 0.0f : s > 65535.0f ? 65535.0f : s;
      v[c] = (uint32_t) rintf(s);
    }
    float s = af * 65535.0f;
    s = s < 0.0f ? 0.0f : s > 65535.0f ? 65535.0f : s;
    alpha = (uint32_t) rintf(s);
  } else {
    const float d = nt.dither[DitherIndex(f, (uint32_t) x, (uint32_t) y)];
    for (int c = 0; c < 3; ++c) v[c] = ToU8Dithered(rgb[c], d);
    alpha = ToU8Dithered(af, d);
  }
  if (!out.alpha_float) alpha = AlphaAt(f, out, x, y);
  if (cp.grey) v[0] = v[2] = v[1];
  StoreRgba(out, x, y, v[0], v[1], v[2], alpha);
}

// 2x upsampling of a frame coded at half resolution (what libjxl's encoder does at distances of about 10 and more): every
// coded XYB sample (x, y) becomes the 2 x 2 output samples (2x + ox, 2y + oy), each a 5 x 5 weighting of the coded
// neighbourhood (borders mirrored) with the default kernel of the format -- a symmetric matrix given by 15 weights,
// mirrored for ox / oy = 1 -- clamped to the range of that neighbourhood.  Unfused multiply / add in row-major order, as
// the reference's SSE2 build accumulates.  src: coded planes (f geometry); dst: [3][up_h][up_stride].
JXLB_HD void StageUpsample2(const FrameDev& f, const float* src, float* dst, uint32_t up_stride, uint32_t up_h, int x, int y) {
  const float kW[15] = {-0.01716200f, -0.03452303f, -0.04022174f, -0.02921014f, -0.00624645f, 0.14111091f, 0.28896755f, 0.00278718f,
                        -0.01610267f, 0.56661550f, 0.03777607f, -0.01986694f, -0.03144731f, -0.01185068f, -0.00213539f};
  const size_t plane = (size_t) f.plane_h * f.plane_stride, uplane = (size_t) up_h * up_stride;
  int xs[5], ys[5];
  for (int k = 0; k < 5; ++k) {
    xs[k] = Mirror(x + k - 2, (int) f.width);
    ys[k] = Mirror(y + k - 2, (int) f.height);
  }
  for (int c = 0; c < 3; ++c) {
    float v[5][5];
    float mn = 0.0f, mx = 0.0f;
    for (int iy = 0; iy < 5; ++iy)
      for (int ix = 0; ix < 5; ++ix) {
        const float t = src[c * plane + (size_t) ys[iy] * f.plane_stride + xs[ix]];
        v[iy][ix] = t;
        if (iy == 0 && ix == 0) mn = mx = t;
        else {
          mn = t < mn ? t : mn;
          mx = t > mx ? t : mx;
        }
      }
    for (int oy = 0; oy < 2; ++oy)
      for (int ox = 0; ox < 2; ++ox) {
        float acc = 0.0f;
        for (int iy = 0; iy < 5; ++iy)
          for (int ix = 0; ix < 5; ++ix) {
            const int ky = oy ? 4 - iy : iy, kx = ox ? 4 - ix : ix;
            const int a = ky < kx ? ky : kx, b = ky < kx ? kx : ky;
            acc = AddRn(MulRn(kW[5 * a - a * (a - 1) / 2 + b - a], v[iy][ix]), acc);
          }
        acc = acc < mn ? mn : acc > mx ? mx : acc;
        dst[c * uplane + (size_t) (2 * y + oy) * up_stride + 2 * x + ox] = acc;
      }
  }
}

// The alpha channel of a half-resolution frame, coded at half resolution too: libjxl turns its integers into floats in
// [0, 1], upsamples them with the same kernel and clamping as the colour planes and converts back on output.  src: the
// coded int32 plane (f.width x f.height, row stride f.mod_stride); dst: [up_h][up_stride] FLOATS in [0, 1] (OutputDesc::
// alpha_float): libjxl quantises them on output like the colour channels -- with the 8-bit dither, which an exact k / 255
// never shows but an upsampled value does.
JXLB_HD void StageUpsampleAlpha2(const FrameDev& f, const int32_t* src, uint32_t bits, int32_t* dst, uint32_t up_stride, int x, int y) {
  const float kW[15] = {-0.01716200f, -0.03452303f, -0.04022174f, -0.02921014f, -0.00624645f, 0.14111091f, 0.28896755f, 0.00278718f,
                        -0.01610267f, 0.56661550f, 0.03777607f, -0.01986694f, -0.03144731f, -0.01185068f, -0.00213539f};
  const float inv = 1.0f / (float) ((1u << bits) - 1);
  float v[5][5];
  float mn = 0.0f, mx = 0.0f;
  for (int iy = 0; iy < 5; ++iy)
    for (int ix = 0; ix < 5; ++ix) {
      const float t = MulRn((float) src[(size_t) Mirror(y + iy - 2, (int) f.height) * f.mod_stride + Mirror(x + ix - 2, (int) f.width)], inv);
      v[iy][ix] = t;
      if (iy == 0 && ix == 0) mn = mx = t;
      else {
        mn = t < mn ? t : mn;
        mx = t > mx ? t : mx;
      }
    }
  for (int oy = 0; oy < 2; ++oy)
    for (int ox = 0; ox < 2; ++ox) {
      float acc = 0.0f;
      for (int iy = 0; iy < 5; ++iy)
        for (int ix = 0; ix < 5; ++ix) {
          const int ky = oy ? 4 - iy : iy, kx = ox ? 4 - ix : ix;
          const int a = ky < kx ? ky : kx, b = ky < kx ? kx : ky;
          acc = AddRn(MulRn(kW[5 * a - a * (a - 1) / 2 + b - a], v[iy][ix]), acc);
        }
      acc = acc < mn ? mn : acc > mx ? mx : acc;
      reinterpret_cast<float*>(dst)[(size_t) (2 * y + oy) * up_stride + 2 * x + ox] = acc;
    }
}

// Modular (non-XYB) frames: integer samples straight to the output depth.
JXLB_HD void StageModularToRgba(const FrameDev& f, const OutputDesc& out, int x, int y) {
  const uint32_t maxout = out.bits16 ? 65535u : 255u;
  const size_t plane = (size_t) f.height * f.mod_stride, o = (size_t) y * f.mod_stride + x;
  uint32_t v[3];
  if (f.num_color_mod_channels == 1) {
    v[0] = v[1] = v[2] = ScaleSample(f.mod[o], out.color_bits, maxout);
  } else {
    for (int c = 0; c < 3; ++c) v[c] = ScaleSample(f.mod[c * plane + o], out.color_bits, maxout);
  }
  StoreRgba(out, x, y, v[0], v[1], v[2], AlphaAt(f, out, x, y));
}

// Frame-level inverse transforms (RCT, palette) of a multi-section frame's modular image, one pixel (the per-group ones
// are undone in the group decoder).  Applied in reverse transform order on the planes PlanChannels assigned (modular.h).
JXLB_HD void StageGlobalInverse(const FrameDev& f, int x, int y) {
  const size_t plane = (size_t) f.height * f.mod_stride, o = (size_t) y * f.mod_stride + x;
  for (int t = (int) f.global_nb_transforms - 1; t >= 0; --t) {
    const ModTransform& tr = f.global_tr[t];
    if (tr.id == 0) {
      int32_t a = f.mod[tr.pl[0] * plane + o], b = f.mod[tr.pl[1] * plane + o], c = f.mod[tr.pl[2] * plane + o];
      InverseRctPixel(tr.rct_type, a, b, c);
      uint32_t perm[3];
      RctPermutation(tr.rct_type, perm);
      f.mod[tr.pl[perm[0]] * plane + o] = a;
      f.mod[tr.pl[perm[1]] * plane + o] = b;
      f.mod[tr.pl[perm[2]] * plane + o] = c;
    } else if (tr.id == 1) {
      int32_t v[kMaxModPlanes];
      // a pixel the palette path cannot reconstruct fails the frame: the status slot of its first group is read back after the run
      if (!InversePalettePixel(tr, f.meta, f.bit_depth, f.mod[tr.pl[0] * plane + o], v)) f.status[f.num_lf_groups] = kErrUnsupported;
      for (uint32_t c = 0; c < tr.num_c; ++c) f.mod[tr.pl[c] * plane + o] = v[c];
    }
  }
}

}  // namespace jxlb
