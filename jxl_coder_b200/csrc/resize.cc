#include "resize.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace jxlb {

namespace {

// pic-scale's BC-spline in f32.  The quantised weights depend on the rounding of every step.  pic-scale's source is not
// available here; of the evaluation orders tried against weights extracted from the reference binary, the expanded
// polynomial fits Mitchell / CatmullRom / Hermite best and the Horner form fits Cubic / BSpline best (see the header
// of resize.h for what remains).
float BcSpline(float d, float b, float c, bool horner) {
  const float x = std::fabs(d);
  const float dp = x * x;
  const float sixth = 1.0f / 6.0f;
  if (x < 1.0f) {
    const float c1 = (12.0f - 9.0f * b) - 6.0f * c;
    const float c2 = (-18.0f + 12.0f * b) + 6.0f * c;
    const float c3 = 6.0f - 2.0f * b;
    if (horner) return ((c1 * x + c2) * dp + c3) * sixth;
    return ((c1 * (dp * x) + c2 * dp) + c3) * sixth;
  }
  if (x < 2.0f) {
    const float c1 = -b - 6.0f * c;
    const float c2 = 6.0f * b + 30.0f * c;
    const float c3 = -12.0f * b - 48.0f * c;
    const float c4 = 8.0f * b + 24.0f * c;
    if (horner) return ((((c1 * x + c2) * x) + c3) * x + c4) * sixth;
    return (((c1 * (dp * x) + c2 * dp) + c3 * x) + c4) * sixth;
  }
  return 0.0f;
}

// pic-scale's Lanczos3: sinc(x) * sinc(x / 3) for |x| < 3, sinc(x) = sin(pi x) / (pi x).  Its sine goes through the pxfm
// crate's f_sinpif (a correctly rounded f32 sinpi); here sin(pi x) is evaluated in double and rounded to f32 once, which
// gives the same f32 wherever the double result is not within 2^-29 of a rounding boundary (bit-exact against the
// reference binary on every size tried, tests/test_resize_host.py).
float Sinc(float x) {
  if (x == 0.0f) return 1.0f;
  const float s = (float) std::sin(M_PI * (double) x);
  return s / (3.14159265358979323846f * x);
}
float Lanczos3(float x) {
  const float a = std::fabs(x);
  if (a >= 3.0f) return 0.0f;
  return Sinc(a) * Sinc(a / 3.0f);
}

struct Kernel {
  int kind;              // 0 bilinear, 1 bc-spline (expanded), 2 bc-spline (Horner), 3 Lanczos3
  float b, c;
  float min_kernel_size;
  float operator()(float x) const {
    if (kind == 0) {
      const float a = std::fabs(x);
      return a < 1.0f ? 1.0f - a : 0.0f;
    }
    if (kind == 3) return Lanczos3(x);
    return BcSpline(x, b, c, kind == 2);
  }
};

bool KernelFor(int32_t filter, Kernel* k) {
  const float third = 1.0f / 3.0f;
  switch (filter) {
    case 1: *k = Kernel{0, 0.f, 0.f, 2.0f}; return true;            // Bilinear
    case 3: *k = Kernel{2, 1.0f, 0.0f, 4.0f}; return true;          // Cubic
    case 8: *k = Kernel{1, 1.0f, 0.0f, 4.0f}; return true;          // BSpline
    case 4: *k = Kernel{1, third, third, 4.0f}; return true;        // MitchellNetravalli
    case 6: *k = Kernel{1, 0.0f, 0.5f, 4.0f}; return true;          // CatmullRom
    case 7: *k = Kernel{1, 0.0f, 0.0f, 4.0f}; return true;          // Hermite
    case 10: *k = Kernel{1, 0.0f, 0.5f, 4.0f}; return true;         // Bicubic: pic-scale 0.7.6 gives CatmullRom's output bit for bit
    case 5: case 9:                                                 // Lanczos3; HANN is mapped to Lanczos (SizeScaler.cpp:86-89)
      *k = Kernel{3, 0.f, 0.f, 6.0f};
      return true;
    default: return false;
  }
}

void MakeAxis(uint32_t in_size, uint32_t out_size, const Kernel& k, ResizeAxis* a) {
  a->in_size = in_size;
  a->out_size = out_size;
  const float scale = (float) in_size / (float) out_size;
  const float cutoff = std::max(scale, 1.0f);
  const uint32_t base_size = (uint32_t) std::floor(k.min_kernel_size * cutoff + 0.5f);
  const float radius = (float) base_size / 2.0f;
  const float fscale = 1.0f / cutoff;
  a->taps = base_size;
  a->start.assign(out_size, 0);
  a->count.assign(out_size, 0);
  a->weights.assign((size_t) out_size * base_size, 0);
  std::vector<float> w(base_size);
  for (uint32_t i = 0; i < out_size; ++i) {
    const float center_x = std::min(((float) i + 0.5f) * scale, (float) in_size);
    const float fs = std::max(std::floor(center_x - radius), 0.0f);
    const uint32_t start = (uint32_t) fs;
    const float fe = std::min(std::min(std::ceil(center_x + radius), (float) (start + base_size)), (float) in_size);
    const uint32_t end = (uint32_t) fe;
    const float center = center_x - 0.5f;
    // f32 tap weights; their sum behaves like an exactly accumulated one (Bilinear, whose taps are exact, pins this and
    // the division below: 0 mismatches over random sizes)
    double sum = 0.0;
    const uint32_t n = end > start ? end - start : 0;
    for (uint32_t t = 0; t < n; ++t) {
      const float dx = std::fabs((float) (start + t) - center);
      w[t] = k(dx * fscale);
      sum += w[t];
    }
    const float fsum = (float) sum;
    a->start[i] = start;
    a->count[i] = n;
    if (fsum != 0.0f) {
      for (uint32_t t = 0; t < n; ++t) {
        const float q = std::trunc((w[t] / fsum) * 32768.0f);
        a->weights[(size_t) i * base_size + t] = (int16_t) std::min(std::max(q, -32768.0f), 32767.0f);
      }
    }
  }
}

void MakeNearestAxis(uint32_t in_size, uint32_t out_size, ResizeAxis* a) {
  a->in_size = in_size;
  a->out_size = out_size;
  a->taps = 1;
  a->start.assign(out_size, 0);
  a->count.assign(out_size, 1);
  a->weights.assign(out_size, 0);
  const uint64_t s = ((uint64_t) in_size << 32) / out_size;
  for (uint32_t i = 0; i < out_size; ++i) a->start[i] = (uint32_t) std::min<uint64_t>(((uint64_t) i * s + (s >> 1)) >> 32, in_size - 1);
}

}  // namespace

int MakeResizePlan(uint32_t src_w, uint32_t src_h, int32_t req_w, int32_t req_h, int32_t scale_mode, int32_t filter, bool has_alpha,
                   ResizePlan* p) {
  if (!src_w || !src_h || req_w == 0 || req_h == 0) return kResizeBadArg;
  if (has_alpha && (filter == 5 || filter == 9)) return kResizeUnsupported;  // see resize.h
  // resolve_dimensions (weaver/src/scale.rs:100-135)
  size_t nw, nh;
  if (req_w > 0 && req_h == -1) {
    const double s = (double) req_w / (double) src_w;
    nw = (size_t) req_w;
    nh = std::max<size_t>((size_t) std::round((double) src_h * s), 1);
  } else if (req_w > 0 && req_h == -2) {
    const double s = (double) req_w / (double) src_w;
    nw = (size_t) req_w;
    nh = (std::max<size_t>((size_t) std::round((double) src_h * s), 1) + 1) & ~(size_t) 1;
  } else if (req_w == -1 && req_h > 0) {
    const double s = (double) req_h / (double) src_h;
    nw = std::max<size_t>((size_t) std::round((double) src_w * s), 1);
    nh = (size_t) req_h;
  } else if (req_w == -2 && req_h > 0) {
    const double s = (double) req_h / (double) src_h;
    nw = (std::max<size_t>((size_t) std::round((double) src_w * s), 1) + 1) & ~(size_t) 1;
    nh = (size_t) req_h;
  } else {
    nw = (size_t) std::max(req_w, 1);
    nh = (size_t) std::max(req_h, 1);
  }
  // scale mode (weaver/src/scale.rs:199-236)
  size_t sw, sh, cx = 0, cy = 0, cw, ch;
  if (scale_mode == 2 || scale_mode == 1) {
    const double xf = (double) nw / (double) src_w, yf = (double) nh / (double) src_h;
    const double s = scale_mode == 2 ? std::max(xf, yf) : std::min(xf, yf);
    sw = std::max<size_t>((size_t) std::round((double) src_w * s), 1);
    sh = std::max<size_t>((size_t) std::round((double) src_h * s), 1);
    cx = (size_t) std::max<int64_t>(((int64_t) sw - (int64_t) nw) / 2, 0);
    cy = (size_t) std::max<int64_t>(((int64_t) sh - (int64_t) nh) / 2, 0);
    cw = std::min(nw, sw);
    ch = std::min(nh, sh);
  } else {
    sw = nw;
    sh = nh;
    cw = nw;
    ch = nh;
  }
  if (sw > 65535 || sh > 65535 || (uint64_t) sw * sh * 4 >= 0x7FFFFFFFull) return kResizeBadArg;
  p->src_w = src_w;
  p->src_h = src_h;
  p->scaled_w = (uint32_t) sw;
  p->scaled_h = (uint32_t) sh;
  p->crop_x = (uint32_t) cx;
  p->crop_y = (uint32_t) cy;
  p->out_w = (uint32_t) cw;
  p->out_h = (uint32_t) ch;
  p->zero_last_row = cx > 0;
  p->identity_v = sh == src_h;
  p->identity_h = sw == src_w;
  p->nearest = filter == 2;
  if (p->nearest) {
    p->identity_v = p->identity_h = false;
    MakeNearestAxis(src_h, (uint32_t) sh, &p->v);
    MakeNearestAxis(src_w, (uint32_t) sw, &p->h);
    return kResizeOk;
  }
  Kernel k;
  if (!KernelFor(filter, &k)) return kResizeUnsupported;
  p->premultiply = has_alpha && !(p->identity_v && p->identity_h);
  if (p->identity_v && !p->identity_h && src_h >= 4) p->zero_tail_rows = src_h % 4;
  if (!p->identity_v) MakeAxis(src_h, (uint32_t) sh, k, &p->v);
  if (!p->identity_h) MakeAxis(src_w, (uint32_t) sw, k, &p->h);
  return kResizeOk;
}

}  // namespace jxlb
