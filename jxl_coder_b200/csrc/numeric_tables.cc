// Host-side construction of the constant tables of the numeric kernels: default dequantisation matrices
// (JPEG XL default quant-weight parameters), LLF synthesis matrices, dither table, AFV basis.
#include "numeric_tables.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

#if defined(__x86_64__) || defined(__i386__)
#include <xmmintrin.h>
#endif

namespace jxlb {
namespace {
// rcp11[i] = rcpps(1 + i / 2048) as the HOST CPU computes it (numeric.h: NumericTables::rcp11).  Checked: on this
// machine RCPPS depends on the top 11 mantissa bits only and scales exactly with the exponent
// (tests/test_numeric_host.py::test_rcp_table_reproduces_rcpps).
// JXLB_EXACT_RCP=1 (or a non-x86 host) fills the table with correctly rounded reciprocals of the bucket's lower edge
// instead: vendor-independent output, at most 1 LSB from the reference's on a few more samples (include/jxlb200.h).
void FillRcp11(uint32_t* out) {
  const char* e = getenv("JXLB_EXACT_RCP");
  const bool exact = e && *e && *e != '0';
  for (uint32_t i = 0; i < 2048; ++i) {
    union { float f; uint32_t u; } v, r;
    v.u = 0x3F800000u | (i << 12);
    r.f = 1.0f / v.f;
#if defined(__x86_64__) || defined(__i386__)
    if (!exact) {
      alignas(16) float in4[4] = {v.f, v.f, v.f, v.f}, out4[4];
      _mm_store_ps(out4, _mm_rcp_ps(_mm_load_ps(in4)));
      r.f = out4[0];
    }
#else
    (void) exact;
#endif
    out[i] = r.u;
  }
}
}  // namespace
}  // namespace jxlb


namespace jxlb {

namespace {

const double kT[7] = {-1.025, -0.78, -0.65012, -0.19041574, -0.208193958, -0.421064, -0.327338457};
const double kU[7] = {-0.304195821, -0.363303632, -0.356603801, -0.344307452, -0.336995929, -0.301808655, -0.273216844};
const double kW[7] = {-1.2, -1.2, -0.8, -0.7, -0.7, -0.4, -0.5};

struct Bands {
  int n;
  double v[3][8];
};

// distance-band parameters of the DCT-type tables, X / Y / B
Bands DctParams(int q) {
  Bands b{};
  auto set = [&](int n, std::initializer_list<double> x, std::initializer_list<double> y, std::initializer_list<double> bb) {
    b.n = n;
    int i = 0;
    for (double v : x) b.v[0][i++] = v;
    i = 0;
    for (double v : y) b.v[1][i++] = v;
    i = 0;
    for (double v : bb) b.v[2][i++] = v;
  };
  auto tail = [&](double x0, double y0, double b0) {
    b.n = 8;
    b.v[0][0] = x0;
    b.v[1][0] = y0;
    b.v[2][0] = b0;
    for (int i = 0; i < 7; ++i) {
      b.v[0][i + 1] = kT[i];
      b.v[1][i + 1] = kU[i];
      b.v[2][i + 1] = kW[i];
    }
  };
  switch (q) {
    case 0: set(6, {3150.0, 0.0, -0.4, -0.4, -0.4, -2.0}, {560.0, 0.0, -0.3, -0.3, -0.3, -0.3}, {512.0, -2.0, -1.0, 0.0, -1.0, -2.0}); break;
    case 3: set(4, {2200.0, 0.0, 0.0, 0.0}, {392.0, 0.0, 0.0, 0.0}, {112.0, -0.25, -0.25, -0.5}); break;
    case 4:
      set(7, {8996.8725711814115328, -1.3000777393353804, -0.49424529824571225, -0.439093774457103443, -0.6350101832695744, -0.90177264050827612, -1.6162099239887414},
          {3191.48366296844234752, -0.67424582104194355, -0.80745813428471001, -0.44925837484843441, -0.35865440981033403, -0.31322389111877305, -0.37615025315725483},
          {1157.50408145487200256, -2.0531423165804414, -1.4, -0.50687130033378396, -0.42708730624733904, -1.4856834539296244, -4.9209142884401604});
      break;
    case 5:
      set(8, {15718.40830982518931456, -1.025, -0.98, -0.9012, -0.4, -0.48819395464, -0.421064, -0.27},
          {7305.7636810695983104, -0.8041958212306401, -0.7633036457487539, -0.55660379990111464, -0.49785304658857626, -0.43699592683512467, -0.40180866526242109, -0.27321683125358037},
          {3803.53173721215041536, -3.060733579805728, -2.0413270132490346, -2.0235650159727417, -0.5495389509954993, -0.4, -0.4, -0.3});
      break;
    case 6: set(7, {7240.7734, -0.7, -0.7, -0.2, -0.2, -0.2, -0.5}, {1448.15466, -0.5, -0.5, -0.5, -0.2, -0.2, -0.2}, {506.854126, -1.4, -0.2, -0.5, -0.5, -1.5, -3.6}); break;
    case 7:
      set(8, {16283.249, -1.78128457, -1.63090587, -1.0382179, -0.85, -0.7, -0.9, -1.23606384}, {5089.15771, -0.320049405, -0.353628486, -0.3034, -0.61, -0.5, -0.5, -0.6},
          {3397.77612, -0.321327358, -0.345076203, -0.7034, -0.9, -1.0, -1.0, -1.17546058});
      break;
    case 8:
      set(8, {13844.9707, -0.971138, -0.658, -0.42026, -0.22712, -0.2206, -0.226, -0.6},
          {4798.96387, -0.611253083, -0.837707877, -0.790148616, -0.269272745, -0.382727683, -0.229242221, -0.20719099},
          {1807.23694, -1.2, -1.2, -0.7, -0.7, -0.7, -0.4, -0.5});
      break;
    case 9: set(4, {2198.05054, -0.962696254, -0.761942506, -0.655114055}, {764.36554, -0.926302016, -0.967522979, -0.278452903}, {527.107544, -1.45943856, -1.45008206, -1.58437228}); break;
    case 11: tail(23966.166, 8380.19141, 4493.02393); break;
    case 12: tail(15358.8984, 5597.36035, 2919.96167); break;
    case 13: tail(47932.332, 16760.3828, 8986.04785); break;
    case 14: tail(30717.7969, 11194.7207, 5839.92334); break;
    case 15: tail(95864.6641, 33520.7656, 17972.0957); break;
    case 16: tail(61435.5938, 22389.4414, 11679.8467); break;
    default: break;
  }
  return b;
}

// libjxl computes its quant weights in float with FastPowf = FastPow2f(FastLog2f(b) * e), two small rational polynomials
// (lib/jxl/base/fast_math-inl.h; L1 error ~4e-6 in the logarithm), not with pow(): the tables differ from the closed form
// by up to ~1e-5 relative, which reaches the 8-bit rounding at higher distances.  Restated operation by operation
// (separately rounded multiplies and adds: the reference's x86_64 library is an SSE2 build); the 13 polynomial constants
// were checked against the shipped lib/x86_64/libjxl.so (each present as a 4-lane vector in .rodata).
float Mult(float v) { return v > 0.0f ? 1.0f + v : 1.0f / (1.0f - v); }

float FastLog2f(float x) {
  int32_t xb;
  memcpy(&xb, &x, 4);
  const int32_t exp_bits = xb - 0x3f2aaaab;   // = 2/3
  const int32_t exp_shifted = exp_bits >> 23;
  const int32_t mb = xb - (int32_t) ((uint32_t) exp_shifted << 23);
  float mantissa;
  memcpy(&mantissa, &mb, 4);
  const float t = mantissa - 1.0f;
  volatile float yp = 7.4245873327820566E-01f, yq = 1.7409343003366853E-01f, m;
  m = yp * t;
  yp = m + 1.4287160470083755E+00f;
  m = yp * t;
  yp = m + -1.8503833400518310E-06f;
  m = yq * t;
  yq = m + 1.0096718572241148E+00f;
  m = yq * t;
  yq = m + 9.9032814277590719E-01f;
  const float ratio = yp / yq;
  return ratio + (float) exp_shifted;
}

float FastPow2f(float x) {
  const float floorx = floorf(x);
  const int32_t eb = (int32_t) ((uint32_t) ((int32_t) floorx + 127) << 23);
  float ex;
  memcpy(&ex, &eb, 4);
  const float frac = x - floorx;
  volatile float num = frac + 1.01749063e+01f, m, den;
  m = num * frac;
  num = m + 4.88687798e+01f;
  m = num * frac;
  num = m + 9.85506591e+01f;
  num = num * ex;
  m = frac * 2.10242958e-01f;
  den = m + -2.22328856e-02f;
  m = den * frac;
  den = m + -1.94414990e+01f;
  m = den * frac;
  den = m + 9.85506633e+01f;
  return num / den;
}

float FastPowf(float base, float exponent) {
  volatile float l = FastLog2f(base);
  volatile float le = l * exponent;
  return FastPow2f(le);
}

// weights of a rows x cols coefficient array from distance bands (App. B.7 "Quant weights"; libjxl GetQuantWeights)
void BandWeights(int rows, int cols, const double* params, int nb, std::vector<double>* w) {
  float bands[17] = {};
  bands[0] = (float) params[0];
  for (int i = 1; i < nb; ++i) bands[i] = bands[i - 1] * Mult((float) params[i]);
  const float scale = (float) (nb - 1) / (1.41421356237f + 1e-6f);
  const float rc = scale / (float) (cols - 1), rr = scale / (float) (rows - 1);
  w->assign((size_t) rows * cols, 0.0);
  for (int y = 0; y < rows; ++y) {
    const float dy = (float) y * rr;
    const float dy2 = dy * dy;
    for (int x = 0; x < cols; ++x) {
      const float dx = (float) x * rc;
      volatile float dx2 = dx * dx;
      const float dist = sqrtf(dx2 + dy2);
      float weight = bands[0];
      if (nb > 1) {
        const int idx = (int) dist;
        const float frac = dist - (float) idx;
        const float a = bands[idx], b = bands[idx + 1];
        volatile float p = FastPowf(b / a, frac);
        weight = a * p;
      }
      (*w)[(size_t) y * cols + x] = weight;
    }
  }
}

// libjxl's scalar Interpolate (AFV table)
float InterpolateBands(float pos, float max, const float* array, int len) {
  const float scaled_pos = pos * (float) (len - 1) / max;
  const int idx = (int) scaled_pos;
  const float a = array[idx], b = array[idx + 1];
  volatile float p = FastPowf(b / a, scaled_pos - (float) idx);
  return a * p;
}

const double kIdWeights[3][3] = {{280.0, 3160.0, 3160.0}, {60.0, 864.0, 864.0}, {18.0, 200.0, 200.0}};
const double kDct2Weights[3][6] = {{3840.0, 2560.0, 1280.0, 640.0, 480.0, 300.0}, {960.0, 640.0, 320.0, 180.0, 140.0, 120.0}, {640.0, 320.0, 128.0, 64.0, 32.0, 16.0}};
const double kAfvWeights[3][9] = {{3072.0, 3072.0, 256.0, 256.0, 256.0, 414.0, 0.0, 0.0, 0.0}, {1024.0, 1024.0, 50.0, 50.0, 50.0, 58.0, 0.0, 0.0, 0.0},
                                  {384.0, 384.0, 12.0, 12.0, 12.0, 22.0, -0.25, -0.25, -0.25}};
const double kAfvFreqs[16] = {0, 0, 0.8517778890324296, 5.37778436506804, 0, 0, 4.734747904497923, 5.449245381693219, 1.6598270267479331, 4,
                              7.275749096817861, 10.423227632456525, 2.662932286148962, 7.630657783650829, 8.962388608184032, 12.97166202570235};

void TableWeights(int q, int c, std::vector<double>* w, int* rows, int* cols) {
  const uint32_t s = QuantTableRepresentative((uint32_t) q);
  const int cx = (int) StrategyCellsX(s), cy = (int) StrategyCellsY(s);
  *rows = 8 * (cx < cy ? cx : cy);
  *cols = 8 * (cx < cy ? cy : cx);
  if (q == 1) {
    w->assign(64, kIdWeights[c][0]);
    (*w)[1] = (*w)[8] = kIdWeights[c][1];
    (*w)[9] = kIdWeights[c][2];
  } else if (q == 2) {
    const double* d = kDct2Weights[c];
    w->assign(64, 0.0);
    auto fill = [&](int y0, int y1, int x0, int x1, double v) {
      for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) (*w)[y * 8 + x] = v;
    };
    fill(0, 4, 4, 8, d[4]);
    fill(4, 8, 0, 4, d[4]);
    fill(4, 8, 4, 8, d[5]);
    fill(0, 2, 2, 4, d[2]);
    fill(2, 4, 0, 2, d[2]);
    fill(2, 4, 2, 4, d[3]);
    (*w)[0] = 1.0;
    (*w)[1] = (*w)[8] = d[0];
    (*w)[9] = d[1];
  } else if (q == 3) {
    Bands b = DctParams(3);
    std::vector<double> w44;
    BandWeights(4, 4, b.v[c], b.n, &w44);
    w->assign(64, 0.0);
    for (int y = 0; y < 8; ++y)
      for (int x = 0; x < 8; ++x) (*w)[y * 8 + x] = w44[(y / 2) * 4 + x / 2];
  } else if (q == 9) {
    Bands b = DctParams(9);
    std::vector<double> w48;
    BandWeights(4, 8, b.v[c], b.n, &w48);
    w->assign(64, 0.0);
    for (int y = 0; y < 8; ++y)
      for (int x = 0; x < 8; ++x) (*w)[y * 8 + x] = w48[(y / 2) * 8 + x];
  } else if (q == 10) {
    const double* a = kAfvWeights[c];
    Bands b48 = DctParams(9), b44 = DctParams(3);
    std::vector<double> w48, w44;
    BandWeights(4, 8, b48.v[c], b48.n, &w48);
    BandWeights(4, 4, b44.v[c], b44.n, &w44);
    const float lo = 0.8517778890324296f, hi = 12.97166202570235f - lo + 1e-6f;
    float bands[4];
    bands[0] = (float) a[5];
    for (int i = 1; i < 4; ++i) bands[i] = bands[i - 1] * Mult((float) a[i + 5]);
    w->assign(64, 0.0);
    (*w)[0] = 1.0;
    (*w)[1 * 8 + 0] = a[0];
    (*w)[0 * 8 + 1] = a[1];
    (*w)[2 * 8 + 0] = a[2];
    (*w)[0 * 8 + 2] = a[3];
    (*w)[2 * 8 + 2] = a[4];
    for (int y = 0; y < 4; ++y)
      for (int x = 0; x < 4; ++x) {
        if (x < 2 && y < 2) continue;
        (*w)[(2 * y) * 8 + 2 * x] = InterpolateBands((float) kAfvFreqs[y * 4 + x] - lo, hi, bands, 4);
      }
    for (int y = 0; y < 4; ++y)
      for (int x = 0; x < 8; ++x) {
        if (x == 0 && y == 0) continue;
        (*w)[(2 * y + 1) * 8 + x] = w48[y * 8 + x];
      }
    for (int y = 0; y < 4; ++y)
      for (int x = 0; x < 4; ++x) {
        if (x == 0 && y == 0) continue;
        (*w)[(2 * y) * 8 + 2 * x + 1] = w44[y * 4 + x];
      }
  } else {
    Bands b = DctParams(q);
    BandWeights(*rows, *cols, b.v[c], b.n, w);
  }
}

const float kDither[1024] = {
#include "tables/dither_table.inc"
};
const float kAfvBasis[256] = {
#include "tables/afv_basis.inc"
};

}  // namespace

const HostNumericTables& GetHostNumericTables() {
  static HostNumericTables t;
  static std::once_flag once;
  std::call_once(once, [] {
    for (int q = 0; q < kNumQuantTables; ++q) {
      bool symmetric = true;
      for (int c = 0; c < 3; ++c) {
        std::vector<double> w;
        int rows, cols;
        TableWeights(q, c, &w, &rows, &cols);
        t.tables.dequant_off[q][c] = (uint32_t) t.dequant_pool.size();
        for (double v : w) t.dequant_pool.push_back(1.0f / (float) v);  // libjxl: table = 1.0f / weight
        if (rows != cols) symmetric = false;
        for (int a = 0; a < rows && symmetric; ++a)
          for (int b = 0; b < a; ++b)
            if ((float) w[(size_t) a * cols + b] != (float) w[(size_t) b * cols + a]) {
              symmetric = false;
              break;
            }
      }
      t.tables.dequant_symmetric[q] = symmetric ? 1 : 0;
    }
    for (int l = 0; l < 6; ++l) {
      const int N = 1 << l;
      t.llf_off[l] = (uint32_t) t.llf_pool.size();
      for (int k = 0; k < N; ++k)
        for (int n = 0; n < N; ++n) {
          const double r = k == 0 ? 1.0 : 8.0 * std::sin(k * M_PI / (16.0 * N)) / std::sin(k * M_PI / (2.0 * N));
          const double ck = k == 0 ? 1.0 : std::sqrt(2.0);
          t.llf_pool.push_back((float) (r * ck / N * std::cos((2 * n + 1) * k * M_PI / (2.0 * N))));
        }
    }
    t.tables.dequant = t.dequant_pool.data();
    for (int l = 0; l < 6; ++l) t.tables.llf[l] = t.llf_pool.data() + t.llf_off[l];
    for (int i = 0; i < 1024; ++i) t.tables.dither[i] = kDither[i];
    FillRcp11(t.tables.rcp11);
    for (int i = 0; i < 256; ++i) t.tables.afv_basis[i] = kAfvBasis[i];
  });
  return t;
}

}  // namespace jxlb
