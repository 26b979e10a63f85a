// Per-sample / per-block arithmetic of the VarDCT reconstruction, host + device (see hd.h): quant-bias adjustment,
// recursive inverse DCT held in registers, the special 8x8 transforms (IDENTITY, DCT2X2, DCT4X4, DCT4X8, AFV),
// Gaborish, the edge-preserving filter and XYB -> RGB with output transfer function and dither.
// The CUDA kernels (kernels_numeric.cu) map threads onto these functions; tests/hostemu runs the same functions in
// plain loops so that their arithmetic is checked against the reference's pixels on a box without a GPU.
// Replaces, for this path, libjxl 0.12.0 behind the reference's DecodeJpegXlOneShot
// (/root/reference/jxlcoder/src/main/cpp/interop/JxlDecoding.cpp:74-175).  Formulas: SURVEY.md App. B.7 / C.
#pragma once
#include <math.h>
#include "frame.h"
#include "tables/dct_consts.h"

namespace jxlb {

static constexpr float kSqrt2f = 1.41421356237f;

// Separately rounded multiply / add / divide / square root: the reference's libjxl is an SSE2 build (build_jxl.sh:107-110),
// whose MulAdd is a multiply followed by an add.  nvcc would contract a * b + c into one FMA.
JXLB_HD float MulRn(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
JXLB_HD float AddRn(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
JXLB_HD float MulAddRn(float a, float b, float c) { return AddRn(MulRn(a, b), c); }
JXLB_HD float SqrtRn(float a) {
#ifdef __CUDA_ARCH__
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}
JXLB_HD float DivRn(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}

// ---- dequantisation ---------------------------------------------------------------------------------------------------
JXLB_HD float AdjustQuantBias(int q, uint32_t c) {
  // biases {X, Y, B} for |q| == 1, and q - 0.145/q beyond (App. B.7 "HF dequant")
  const float b = c == 0 ? 0.945349932f : c == 1 ? 0.929945469f : 0.950064898f;
  if (q == 0) return 0.0f;
  if (q == 1) return b;
  if (q == -1) return -b;
  const float f = (float) q;
  return f - 0.145f / f;
}
// The same with libjxl's arithmetic (JXL_HIGH_PRECISION=0 build): q - 0.145 * ApproximateReciprocal(q), the reciprocal
// being x86 RCPPS (rcp11: see NumericTables), multiply and subtract rounded separately.
JXLB_HD float AdjustQuantBiasRcp(int q, uint32_t c, const uint32_t* rcp11) {
  const float b = c == 0 ? 0.945349932f : c == 1 ? 0.929945469f : 0.950064898f;
  if (q == 0) return 0.0f;
  if (q == 1) return b;
  if (q == -1) return -b;
  const float f = (float) q;
  union { float f; uint32_t u; } v;
  v.f = fabsf(f);
  const uint32_t e = (v.u >> 23) & 0xFFu;
  v.u = rcp11[(v.u >> 12) & 0x7FFu] - ((e - 127u) << 23);
  const float r = q < 0 ? -v.f : v.f;
#ifdef __CUDA_ARCH__
  return __fsub_rn(f, __fmul_rn(0.145f, r));
#else
  return f - 0.145f * r;
#endif
}

// ---- inverse DCT: y[n] = sum_k c_k X[k] cos((2n+1) k pi / 2N), c_0 = 1, c_k = sqrt(2) ---------------------------------
// Recursive even/odd decomposition; with N a compile-time constant everything unrolls into registers.
template <int N>
JXLB_HD void Idct1d(float (&v)[N]) {
  if constexpr (N == 1) {
    return;
  } else if constexpr (N == 2) {
    const float a = v[0], b = v[1];
    v[0] = a + b;
    v[1] = a - b;
  } else {
    float e[N / 2], o[N / 2];
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      e[i] = v[2 * i];
      o[i] = v[2 * i + 1];
    }
    Idct1d<N / 2>(e);
#pragma unroll
    for (int i = N / 2 - 1; i > 0; --i) o[i] += o[i - 1];
    o[0] *= kSqrt2f;
    Idct1d<N / 2>(o);
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const float m = WcMul<N>(i) * o[i];
      v[i] = e[i] + m;
      v[N - 1 - i] = e[i] - m;
    }
  }
}

// In-place 1-D inverse DCT of `n` samples at p[0], p[stride], ... (n in {1,2,4,...,64}).
template <int N>
JXLB_HD void Idct1dStrided(float* p, int stride) {
  float v[N];
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = p[i * stride];
  Idct1d<N>(v);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i * stride] = v[i];
}
JXLB_HD void Idct1dDispatch(float* p, int stride, int n) {
  switch (n) {
    case 2: Idct1dStrided<2>(p, stride); break;
    case 4: Idct1dStrided<4>(p, stride); break;
    case 8: Idct1dStrided<8>(p, stride); break;
    case 16: Idct1dStrided<16>(p, stride); break;
    case 32: Idct1dStrided<32>(p, stride); break;
    case 64: Idct1dStrided<64>(p, stride); break;
    default: break;
  }
}

// Memory-based variant for the rare 128 / 256 transforms (too large for registers): same recursion on a scratch
// array of 2n floats.
JXLB_HD_NOINLINE void Idct1dLarge(float* p, int stride, int n, float* scratch) {
  // iterative formulation of the recursion: bit-reversal-like even/odd split done level by level
  // level 0: gather into scratch[0..n)
  float* a = scratch;
  float* b = scratch + n;
  for (int i = 0; i < n; ++i) a[i] = p[i * stride];
  // Recursive helper via explicit stack is overkill; since n <= 256 we unroll the top levels down to 64 and use the
  // register version below that.
  // split once (n -> 2 x n/2), possibly twice (256 -> 4 x 64)
  const int levels = n == 256 ? 2 : 1;
  // forward even/odd splits
  int seg = n;
  for (int l = 0; l < levels; ++l) {
    for (int s = 0; s < n; s += seg) {
      for (int i = 0; i < seg / 2; ++i) {
        b[s + i] = a[s + 2 * i];
        b[s + seg / 2 + i] = a[s + 2 * i + 1];
      }
    }
    float* t = a;
    a = b;
    b = t;
    seg /= 2;
  }
  // seg == 64: odd segments need the B^T step *before* their transform, at each level where they were the odd half.
  // Process recursively in the right order: apply B^T to odd halves top-down.
  // For levels == 1: segments [0,64) even, [64,128) odd (of the 128 transform).
  // For levels == 2: [0,64) ee, [64,128) eo, [128,192) oe, [192,256) oo.
  if (levels == 1) {
    for (int i = 63; i > 0; --i) a[64 + i] += a[64 + i - 1];
    a[64] *= kSqrt2f;
    Idct1dStrided<64>(a, 1);
    Idct1dStrided<64>(a + 64, 1);
    for (int i = 0; i < 64; ++i) {
      const float m = WcMul<128>(i) * a[64 + i];
      p[i * stride] = a[i] + m;
      p[(127 - i) * stride] = a[i] - m;
    }
  } else {
    // undo the second split for the odd half: the B^T of the 256-level acts on the *unsplit* odd half (128 samples).
    // Rebuild that half in natural order, apply B^T, split again.
    for (int i = 0; i < 64; ++i) {
      b[128 + 2 * i] = a[128 + i];
      b[128 + 2 * i + 1] = a[192 + i];
    }
    for (int i = 127; i > 0; --i) b[128 + i] += b[128 + i - 1];
    b[128] *= kSqrt2f;
    for (int i = 0; i < 64; ++i) {
      a[128 + i] = b[128 + 2 * i];
      a[192 + i] = b[128 + 2 * i + 1];
    }
    // each 128-half: odd quarter gets its own B^T, then 64-point transforms, then the 128-level butterfly
    for (int h = 0; h < 2; ++h) {
      float* q = a + 128 * h;
      for (int i = 63; i > 0; --i) q[64 + i] += q[64 + i - 1];
      q[64] *= kSqrt2f;
      Idct1dStrided<64>(q, 1);
      Idct1dStrided<64>(q + 64, 1);
      for (int i = 0; i < 64; ++i) {
        const float m = WcMul<128>(i) * q[64 + i];
        b[128 * h + i] = q[i] + m;
        b[128 * h + 127 - i] = q[i] - m;
      }
    }
    for (int i = 0; i < 128; ++i) {
      const float m = WcMul<256>(i) * b[128 + i];
      p[i * stride] = b[i] + m;
      p[(255 - i) * stride] = b[i] - m;
    }
  }
}

// 2-D inverse DCT of an R x C block stored as F[v][u] at t[v * stride + u] (any sizes up to 64), in place.
JXLB_HD void Idct2dInPlace(float* t, int stride, int R, int C) {
  for (int u = 0; u < C; ++u) Idct1dDispatch(t + u, stride, R);
  for (int v = 0; v < R; ++v) Idct1dDispatch(t + v * stride, 1, C);
}

// ---- special 8x8 transforms -------------------------------------------------------------------------------------------
// `rect` is the block's 8x8 tile in plane layout, i.e. the coefficient array K TRANSPOSED (rect[c][r] = K[r][c]),
// with the LF value already placed at rect[0]; it is replaced by the 8x8 pixels.
JXLB_HD void Idct2Top(const float* in, float* out, int S) {
  const int n = S / 2;
  for (int y = 0; y < n; ++y)
    for (int x = 0; x < n; ++x) {
      const float c00 = in[y * 8 + x], c01 = in[y * 8 + n + x], c10 = in[(y + n) * 8 + x], c11 = in[(y + n) * 8 + n + x];
      out[2 * y * 8 + 2 * x] = c00 + c01 + c10 + c11;
      out[2 * y * 8 + 2 * x + 1] = c00 + c01 - c10 - c11;
      out[(2 * y + 1) * 8 + 2 * x] = c00 - c01 + c10 - c11;
      out[(2 * y + 1) * 8 + 2 * x + 1] = c00 - c01 - c10 + c11;
    }
}

JXLB_HD_NOINLINE void SpecialTransform8x8(uint32_t strategy, float* rect, int stride, const float* afv_basis) {
  float K[64], px[64];
  for (int r = 0; r < 8; ++r)
    for (int c = 0; c < 8; ++c) K[r * 8 + c] = rect[c * stride + r];
  if (strategy == 2) {  // DCT2X2: three levels of 2x2 Hadamard synthesis
    for (int i = 0; i < 64; ++i) px[i] = K[i];
    Idct2Top(K, px, 2);
    for (int i = 0; i < 64; ++i) K[i] = px[i];
    Idct2Top(K, px, 4);
    for (int i = 0; i < 64; ++i) K[i] = px[i];
    Idct2Top(K, px, 8);
  } else if (strategy == 1 || strategy == 3) {  // IDENTITY / DCT4X4
    const float b00 = K[0], b01 = K[1], b10 = K[8], b11 = K[9];
    const float dcs[4] = {b00 + b01 + b10 + b11, b00 + b01 - b10 - b11, b00 - b01 + b10 - b11, b00 - b01 - b10 + b11};
    for (int y = 0; y < 2; ++y)
      for (int x = 0; x < 2; ++x) {
        float sub[16];
        for (int iy = 0; iy < 4; ++iy)
          for (int ix = 0; ix < 4; ++ix) sub[iy * 4 + ix] = K[(y + iy * 2) * 8 + x + ix * 2];
        if (strategy == 3) {
          sub[0] = dcs[y * 2 + x];
          // square: F[v][u] = sub[u][v]
          float F[16];
          for (int v = 0; v < 4; ++v)
            for (int u = 0; u < 4; ++u) F[v * 4 + u] = sub[u * 4 + v];
          for (int u = 0; u < 4; ++u) Idct1dStrided<4>(F + u, 4);
          for (int v = 0; v < 4; ++v) Idct1dStrided<4>(F + v * 4, 1);
          for (int iy = 0; iy < 4; ++iy)
            for (int ix = 0; ix < 4; ++ix) px[(4 * y + iy) * 8 + 4 * x + ix] = F[iy * 4 + ix];
        } else {
          float resid = 0.0f;
          for (int i = 1; i < 16; ++i) resid += sub[i];
          const float base = dcs[y * 2 + x] - resid * (1.0f / 16.0f);
          for (int iy = 0; iy < 4; ++iy)
            for (int ix = 0; ix < 4; ++ix) px[(4 * y + iy) * 8 + 4 * x + ix] = sub[iy * 4 + ix] + base;
          px[(4 * y + 1) * 8 + 4 * x + 1] = base;
          px[(4 * y) * 8 + 4 * x] = sub[5] + base;
        }
      }
  } else if (strategy == 12 || strategy == 13) {  // DCT4X8 / DCT8X4
    const float b0 = K[0], b1 = K[8];
    const float dd[2] = {b0 + b1, b0 - b1};
    for (int s = 0; s < 2; ++s) {
      float F[32];
      for (int iy = 0; iy < 4; ++iy)
        for (int ix = 0; ix < 8; ++ix) F[iy * 8 + ix] = K[(s + iy * 2) * 8 + ix];
      F[0] = dd[s];
      for (int u = 0; u < 8; ++u) Idct1dStrided<4>(F + u, 8);
      for (int v = 0; v < 4; ++v) Idct1dStrided<8>(F + v * 8, 1);
      for (int iy = 0; iy < 4; ++iy)
        for (int ix = 0; ix < 8; ++ix) {
          if (strategy == 12) px[(s * 4 + iy) * 8 + ix] = F[iy * 8 + ix];
          else px[ix * 8 + s * 4 + iy] = F[iy * 8 + ix];
        }
    }
  } else {  // AFV0..3 (14..17)
    const uint32_t kind = strategy - 14;
    const int ax = kind & 1, ay = kind >> 1;
    const float b00 = K[0], b01 = K[1], b10 = K[8];
    const float dcs[3] = {(b00 + b10 + b01) * 4.0f, b00 + b10 - b01, b00 - b10};
    float co[16];
    for (int iy = 0; iy < 4; ++iy)
      for (int ix = 0; ix < 4; ++ix) co[iy * 4 + ix] = K[iy * 2 * 8 + ix * 2];
    co[0] = dcs[0];
    float blk[16];
    for (int i = 0; i < 16; ++i) {
      float acc = 0.0f;
      for (int j = 0; j < 16; ++j) acc += co[j] * afv_basis[j * 16 + i];
      blk[i] = acc;
    }
    for (int iy = 0; iy < 4; ++iy)
      for (int ix = 0; ix < 4; ++ix) px[(iy + ay * 4) * 8 + ax * 4 + ix] = blk[(ay ? 3 - iy : iy) * 4 + (ax ? 3 - ix : ix)];
    {
      float sub[16], F[16];
      for (int iy = 0; iy < 4; ++iy)
        for (int ix = 0; ix < 4; ++ix) sub[iy * 4 + ix] = K[iy * 2 * 8 + ix * 2 + 1];
      sub[0] = dcs[1];
      for (int v = 0; v < 4; ++v)
        for (int u = 0; u < 4; ++u) F[v * 4 + u] = sub[u * 4 + v];
      for (int u = 0; u < 4; ++u) Idct1dStrided<4>(F + u, 4);
      for (int v = 0; v < 4; ++v) Idct1dStrided<4>(F + v * 4, 1);
      const int x0 = ax ? 0 : 4;
      for (int iy = 0; iy < 4; ++iy)
        for (int ix = 0; ix < 4; ++ix) px[(ay * 4 + iy) * 8 + x0 + ix] = F[iy * 4 + ix];
    }
    {
      float F[32];
      for (int iy = 0; iy < 4; ++iy)
        for (int ix = 0; ix < 8; ++ix) F[iy * 8 + ix] = K[(1 + iy * 2) * 8 + ix];
      F[0] = dcs[2];
      for (int u = 0; u < 8; ++u) Idct1dStrided<4>(F + u, 8);
      for (int v = 0; v < 4; ++v) Idct1dStrided<8>(F + v * 8, 1);
      const int y0 = ay ? 0 : 4;
      for (int iy = 0; iy < 4; ++iy)
        for (int ix = 0; ix < 8; ++ix) px[(y0 + iy) * 8 + ix] = F[iy * 8 + ix];
    }
  }
  for (int r = 0; r < 8; ++r)
    for (int c = 0; c < 8; ++c) rect[r * stride + c] = px[r * 8 + c];
}

JXLB_HD bool IsSpecial8x8(uint32_t s) { return s == 1 || s == 2 || s == 3 || (s >= 12 && s <= 17); }

// ---- constant tables used by the numeric kernels (host-built, one copy in HBM) --------------------------------------
struct NumericTables {
  const float* dequant;                              // pool of 1/weight matrices
  uint32_t dequant_off[kNumQuantTables][3];          // per quant table and channel (X, Y, B)
  uint8_t dequant_symmetric[kNumQuantTables];        // square table with w[a][b] == w[b][a] in all three channels
  uint8_t pad_[3];
  // LLF synthesis: A_N[k][n] = r(N, k) * (c_k / N) * cos((2n+1) k pi / 2N), N = 1, 2, 4, 8, 16, 32 at llf[log2 N]
  const float* llf[6];
  alignas(16) float dither[1024];                                // 32 x 32 (App. B.7)
  float afv_basis[256];
  // The reference's libjxl is built with JXL_HIGH_PRECISION=0 (build_jxl.sh:36-37): its EPF normalises with
  // ApproximateReciprocal, i.e. x86 RCPPS -- an 11-bit table lookup on the mantissa whose values are the CPU vendor's.
  // rcp11[i] = bit pattern of rcpps(1 + i / 2048), filled from the host CPU's own instruction (numeric_tables.cc), so the
  // pictures match what the reference produces on the same machine.
  uint32_t rcp11[2048];
};

// rcpps(w) for a positive normal w, from the table above: the result depends on the top 11 mantissa bits only and scales
// exactly with the exponent.
JXLB_HD float ApproxRcp(const uint32_t* rcp11, float w) {
  union { float f; uint32_t u; } v;
  v.f = w;
  const uint32_t e = (v.u >> 23) & 0xFFu;
  v.u = rcp11[(v.u >> 12) & 0x7FFu] - ((e - 127u) << 23);
  return v.f;
}

JXLB_HD int Mirror(int i, int n) {
  // symmetric extension without repeating the edge sample twice: -1 -> 0, -2 -> 1, n -> n-1
  while (i < 0 || i >= n) {
    if (i < 0) i = -i - 1;
    if (i >= n) i = 2 * n - 1 - i;
  }
  return i;
}

// Planar 3-channel float image view with mirrored borders.
struct Planes3 {
  const float* p[3];
  int w, h, stride;
  JXLB_HD float at(int c, int x, int y) const { return p[c][(size_t) Mirror(y, h) * stride + Mirror(x, w)]; }
};

// Gaborish: 3x3 smoothing-undo kernel (App. B.7), one output sample.
template <class Img>
JXLB_HD float GaborishSample(const Img& im, int c, int x, int y, float w1, float w2) {
  const float centre = im.at(c, x, y);
  const float cross = im.at(c, x, y - 1) + im.at(c, x, y + 1) + im.at(c, x - 1, y) + im.at(c, x + 1, y);
  const float diag = im.at(c, x - 1, y - 1) + im.at(c, x + 1, y - 1) + im.at(c, x - 1, y + 1) + im.at(c, x + 1, y + 1);
  const float norm = 1.0f / (1.0f + 4.0f * (w1 + w2));
  return (centre + w1 * cross + w2 * diag) * norm;
}

// Per-cell 1/sigma of the edge-preserving filter (negative; <= kEpfSkip means "leave the cell unfiltered").
static constexpr float kEpfSkipThreshold = -3.90524291751269967f;
JXLB_HD float EpfInvSigma(const FrameDev& f, uint32_t hf_mul, uint32_t sharp) {
  const float qs = (float) f.global_scale * (1.0f / 65536.0f);
  float sigma = f.rf.epf_quant_mul / (qs * (float) hf_mul * -1.1715728752538099f) * f.rf.epf_sharp_lut[sharp];
  if (!(sigma < -1e-4f)) sigma = -1e-4f;
  return 1.0f / sigma;
}

// One EPF output pixel for stage kStage = 0 / 1 / 2 (App. B.7).  Neighbour lists are compile-time constants so that the
// loops unroll completely and every tap becomes a fixed-offset load.
// (px, py) = image coordinates deciding the 8x8-border weighting; (x, y) = coordinates inside `im`.
template <int kStage, class Img>
JXLB_HD void EpfPixelT(const Img& im, const RestorationFilter& rf, const uint32_t* rcp11, int x, int y, int px, int py, float inv_sigma,
                       float out[3]) {
  const float c0 = im.at(0, x, y), c1 = im.at(1, x, y), c2 = im.at(2, x, y);
  if (inv_sigma < kEpfSkipThreshold) {
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    return;
  }
  float sm = 1.65f;
  if (kStage == 0) sm *= rf.epf_pass0_sigma_scale;
  if (kStage == 2) sm *= rf.epf_pass2_sigma_scale;
  const int xm = px & 7, ym = py & 7;
  if (xm == 0 || xm == 7 || ym == 0 || ym == 7) sm *= rf.epf_border_sad_mul;
  const float isg = inv_sigma * sm;
  float wsum = 1.0f, a0 = c0, a1 = c1, a2 = c2;
  constexpr int kN = kStage == 0 ? 12 : 4;
  constexpr int kDx[12] = {0, -1, 0, 1, -2, -1, 1, 2, -1, 0, 1, 0};
  constexpr int kDy[12] = {-2, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 2};
  constexpr int kDx4[4] = {0, 0, -1, 1};
  constexpr int kDy4[4] = {-1, 1, 0, 0};
  // centre plus-shape, reused by every neighbour
  float cc[3][5];
  if (kStage != 2) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      cc[c][0] = c == 0 ? c0 : c == 1 ? c1 : c2;
      cc[c][1] = im.at(c, x, y - 1);
      cc[c][2] = im.at(c, x, y + 1);
      cc[c][3] = im.at(c, x - 1, y);
      cc[c][4] = im.at(c, x + 1, y);
    }
  }
#pragma unroll
  for (int i = 0; i < kN; ++i) {
    const int dx = kStage == 0 ? kDx[i] : kDx4[i];
    const int dy = kStage == 0 ? kDy[i] : kDy4[i];
    float sad = 0.0f;
    float nv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float sc = rf.epf_channel_scale[c];
      nv[c] = im.at(c, x + dx, y + dy);
      if (kStage == 2) {
        sad += fabsf(nv[c] - (c == 0 ? c0 : c == 1 ? c1 : c2)) * sc;
      } else {
        float s = fabsf(nv[c] - cc[c][0]);
        s += fabsf(im.at(c, x + dx, y + dy - 1) - cc[c][1]);
        s += fabsf(im.at(c, x + dx, y + dy + 1) - cc[c][2]);
        s += fabsf(im.at(c, x + dx - 1, y + dy) - cc[c][3]);
        s += fabsf(im.at(c, x + dx + 1, y + dy) - cc[c][4]);
        sad += s * sc;
      }
    }
    float w = 1.0f + sad * isg;
    if (w < 0.0f) w = 0.0f;
    wsum += w;
    a0 += w * nv[0];
    a1 += w * nv[1];
    a2 += w * nv[2];
  }
  const float inv = ApproxRcp(rcp11, wsum);  // libjxl's ApproximateReciprocal (JXL_HIGH_PRECISION=0)
  out[0] = a0 * inv;
  out[1] = a1 * inv;
  out[2] = a2 * inv;
}

template <class Img>
JXLB_HD void EpfPixel(const Img& im, const RestorationFilter& rf, const uint32_t* rcp11, int stage, int x, int y, int px, int py,
                      float inv_sigma, float out[3]) {
  if (stage == 0) EpfPixelT<0>(im, rf, rcp11, x, y, px, py, inv_sigma, out);
  else if (stage == 1) EpfPixelT<1>(im, rf, rcp11, x, y, px, py, inv_sigma, out);
  else EpfPixelT<2>(im, rf, rcp11, x, y, px, py, inv_sigma, out);
}

// ---- colour ------------------------------------------------------------------------------------------------------------
struct ColorParams {
  float opsin_inv[9];      // (linear sRGB -> target primaries) x inverse opsin matrix, scaled by 255 / intensity_target
  float to_target[9];      // linear sRGB -> linear target primaries (identity for sRGB); folded into opsin_inv
  uint32_t apply_primaries;
  uint32_t transfer;       // jxl/color_encoding.h JxlTransferFunction; 0xFFFF = gamma
  float gamma;             // exponent for have_gamma (encoded = linear ^ gamma)
  float pq_scale;          // intensity_target / 10000 for PQ output
  uint32_t grey;           // output is a grey image (R = G = B = luma channel)
};

// v ^ e for v in (0, 1]: on the device through the MUFU log2 / exp2 units (absolute error of lg2.approx is ~2^-22, so
// the result is within ~3e-7 relative of powf -- 1e-4 of an 8-bit LSB -- at a tenth of the instructions).
JXLB_HD float PowUnit(float v, float e) {
#ifdef __CUDA_ARCH__
  return exp2f(__log2f(v) * e);
#else
  return powf(v, e);
#endif
}
// sRGB encode as the reference's libjxl computes it: the library is a JXL_HIGH_PRECISION=0 build (build_jxl.sh:36-37), whose
// linear -> sRGB stage is FastLinearToSRGB, not the closed form: the mantissa is moved to [0.25, 0.5) and run through a
// cubic, the exponent selects one of 16 multipliers 2 * 2^(5 e / 12) kept with a 13-bit mantissa, then * and - 0.055 are
// rounded separately.  It is ~1e-4 away from 1.055 v^(1/2.4) - 0.055 -- enough to flip the 8-bit rounding of ~0.5 % of
// the samples, which was the whole difference between this decoder and the reference on sRGB pictures.  Constants and
// tables read out of the shipped lib/x86_64/libjxl.so (function at 0x2371c0: tables at .rodata 0x11f10 / 0x130f0).
JXLB_HD float SrgbOetf(float v) {
  if (v < 0.0031308f) return MulRn(v, 12.92f);
  union { float f; uint32_t u; } in, m, mul;
  in.f = v;
  m.u = (in.u & 0x007FFFFFu) | 0x3E800000u;
  const float x = m.f;
  float d = MulAddRn(x, 0.059914046f, -0.108894556f);
  d = MulAddRn(d, x, 0.107963754f);
  d = MulAddRn(d, x, 0.018092343f);
  // multipliers for exponents 118 .. 133 (v in [2^-9, 2^7)): bits 25-18 and 17-10 from two 16-entry byte tables
  // {00 0a 19 26 32 41 4d 5c 68 75 83 8f a0 aa b9 c6} and {00 b7 04 0d cb e7 41 68 51 d1 eb f2 00 b7 04 0d}
  const uint32_t e = ((in.u >> 23) - 118u) & 15u;
#ifdef __CUDA_ARCH__
  // byte-permute lookups (the tables live in immediates; an indexed local array would go through local memory)
  const uint32_t hi = e < 8 ? __byte_perm(0x26190a00u, 0x5c4d4132u, e) : __byte_perm(0x8f837568u, 0xc6b9aaa0u, e - 8);
  const uint32_t lo = e < 8 ? __byte_perm(0x0d04b700u, 0x6841e7cbu, e) : __byte_perm(0xf2ebd151u, 0x0d04b700u, e - 8);
  mul.u = ((hi & 0xFFu) << 18) | ((lo & 0xFFu) << 10) | 0x40000000u;
#else
  static const uint8_t kHi[16] = {0x00, 0x0a, 0x19, 0x26, 0x32, 0x41, 0x4d, 0x5c, 0x68, 0x75, 0x83, 0x8f, 0xa0, 0xaa, 0xb9, 0xc6};
  static const uint8_t kLo[16] = {0x00, 0xb7, 0x04, 0x0d, 0xcb, 0xe7, 0x41, 0x68, 0x51, 0xd1, 0xeb, 0xf2, 0x00, 0xb7, 0x04, 0x0d};
  mul.u = ((uint32_t) kHi[e] << 18) | ((uint32_t) kLo[e] << 10) | 0x40000000u;
#endif
  return MulAddRn(d, mul.f, -0.055f);
}
JXLB_HD float Rec709Oetf(float v) {
  if (v < 0.018f) return 4.5f * v;
  return 1.099f * powf(v, 0.45f) - 0.099f;
}
// PQ (SMPTE ST 2084) encode as libjxl 0.12.0 computes it for Rec.2100-PQ tagged pictures (TF_PQ::EncodedFromDisplay):
// NOT the closed form with two powf calls, but a 4-over-4 rational polynomial in x^(1/4) (two square roots), one
// coefficient set below 1e-4 and one above, Horner's scheme with separately rounded multiplies and adds, a true
// division; the sign of the input is carried through.  `v` = linear light already scaled so that 1.0 = 10000 nits.
// The 20 coefficients were read out of the reference's shipped lib/x86_64/libjxl.so (.rodata, each replicated four
// times for the SSE2 lanes) -- libjxl's source is not part of /root/reference.  PQ's slope at black is ~10^6 code values
// per unit of linear light, so the closed form and the approximation (max error 3e-6) disagree by many codes there.
JXLB_HD float PqOetf(float v) {
  const float a = fabsf(v);
  const float x = SqrtRn(SqrtRn(a));
  float yp, yq;
  if (a < 1e-4f) {
    yp = MulAddRn(-2.864824e+05f, x, 6.889862e+04f);
    yp = MulAddRn(yp, x, 1.352821e+02f);
    yp = MulAddRn(yp, x, 3.881234e-01f);
    yp = MulAddRn(yp, x, 9.863406e-06f);
    yq = MulAddRn(-2.072546e+05f, x, -4.389884e+04f);
    yq = MulAddRn(yq, x, 1.608477e+04f);
    yq = MulAddRn(yq, x, 1.477719e+03f);
    yq = MulAddRn(yq, x, 3.371868e+01f);
  } else {
    yp = MulAddRn(4.838434e+01f, x, 1.492516e+02f);
    yp = MulAddRn(yp, x, 5.522776e+01f);
    yp = MulAddRn(yp, x, -1.095778e+00f);
    yp = MulAddRn(yp, x, 1.351392e-02f);
    yq = MulAddRn(2.590418e+01f, x, 1.120607e+02f);
    yq = MulAddRn(yq, x, 9.263710e+01f);
    yq = MulAddRn(yq, x, 2.016708e+01f);
    yq = MulAddRn(yq, x, 1.012416e+00f);
  }
  const float m = DivRn(yp, yq);
  return v < 0.0f ? -m : m;
}

JXLB_HD void XybToEncodedRgb(float x, float y, float b, const ColorParams& cp, float rgb[3]) {
  // libjxl's XybToRgb, operation by operation, with the separately rounded multiplies and adds of its SSE2 build:
  // gamma = (y +- x) - cbrt(-bias); mixed = gamma^2 * gamma + (-bias); linear = M[.][0] * mr, then + M[.][1] * mg, then
  // + M[.][2] * mb.  cp.opsin_inv already contains the primaries conversion and the 255 / intensity_target scale.
  const float kBias = 0.0037930732552754493f;
  const float kCbrtBias = 0.15595420054f;  // cbrt(kBias)
  const float gr = AddRn(AddRn(y, x), kCbrtBias), gg = AddRn(AddRn(y, -x), kCbrtBias), gb = AddRn(b, kCbrtBias);
  const float mr = MulAddRn(MulRn(gr, gr), gr, -kBias), mg = MulAddRn(MulRn(gg, gg), gg, -kBias), mb = MulAddRn(MulRn(gb, gb), gb, -kBias);
  float lin[3];
  for (int i = 0; i < 3; ++i)
    lin[i] = MulAddRn(cp.opsin_inv[3 * i + 2], mb, MulAddRn(cp.opsin_inv[3 * i + 1], mg, MulRn(cp.opsin_inv[3 * i], mr)));
  for (int i = 0; i < 3; ++i) {
    float v = lin[i];
    if (cp.transfer == 16) {
      v = PqOetf(v * cp.pq_scale);
    } else {
      v = v < 0.0f ? 0.0f : v > 1.0f ? 1.0f : v;
      if (cp.transfer == 13) v = SrgbOetf(v);
      else if (cp.transfer == 1) v = Rec709Oetf(v);
      else if (cp.transfer == 17) v = powf(v, 1.0f / 2.6f);
      else if (cp.transfer == 0xFFFF) v = powf(v, cp.gamma);
      // 8 (linear): identity
    }
    rgb[i] = v;
  }
}

JXLB_HD uint8_t ToU8Dithered(float v, float dither) {
  float s = v * 255.0f + dither;
  s = s < 0.0f ? 0.0f : s > 255.0f ? 255.0f : s;
  return (uint8_t) rintf(s);
}

}  // namespace jxlb
