// Host-side construction of the constant tables of the numeric kernels: default dequantisation matrices
// (JPEG XL default quant-weight parameters), LLF synthesis matrices, dither table, AFV basis.
#include "numeric_tables.h"

#include <cmath>
#include <mutex>

#if defined(__x86_64__) || defined(__i386__)
#include <xmmintrin.h>
#endif

namespace jxlb {
namespace {
// rcp11[i] = rcpps(1 + i / 2048) as the HOST CPU computes it (numeric.h: NumericTables::rcp11).  Checked: on this
// machine RCPPS depends on the top 11 mantissa bits only and scales exactly with the exponent
// (tests/test_numeric_host.py::test_rcp_table_reproduces_rcpps).
void FillRcp11(uint32_t* out) {
  for (uint32_t i = 0; i < 2048; ++i) {
    union { float f; uint32_t u; } v, r;
    v.u = 0x3F800000u | (i << 12);
#if defined(__x86_64__) || defined(__i386__)
    alignas(16) float in4[4] = {v.f, v.f, v.f, v.f}, out4[4];
    _mm_store_ps(out4, _mm_rcp_ps(_mm_load_ps(in4)));
    r.f = out4[0];
#else
    r.f = 1.0f / v.f;
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

double Mult(double v) { return v > 0 ? 1 + v : 1 / (1 - v); }

// weights of a rows x cols coefficient array from distance bands (App. B.7 "Quant weights")
void BandWeights(int rows, int cols, const double* params, int nb, std::vector<double>* w) {
  double bands[8];
  bands[0] = params[0];
  for (int i = 1; i < nb; ++i) bands[i] = bands[i - 1] * Mult(params[i]);
  const double scale = (nb - 1) / (std::sqrt(2.0) + 1e-6);
  const double rc = scale / (cols - 1), rr = scale / (rows - 1);
  w->assign((size_t) rows * cols, 0.0);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      const double d = std::hypot(x * rc, y * rr);
      int i = (int) d;
      if (i > nb - 1) i = nb - 1;
      const double fr = d - i;
      const double a = bands[i], b = bands[i + 1 < nb ? i + 1 : nb - 1];
      (*w)[(size_t) y * cols + x] = a * std::pow(b / a, fr);
    }
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
    const double lo = 0.8517778890324296, hi = 12.97166202570235 - lo + 1e-6;
    double bands[4];
    bands[0] = a[5];
    for (int i = 1; i < 4; ++i) bands[i] = bands[i - 1] * Mult(a[i + 5]);
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
        const double pos = (kAfvFreqs[y * 4 + x] - lo) * 3 / hi;
        int i = (int) pos;
        if (i > 2) i = 2;
        (*w)[(2 * y) * 8 + 2 * x] = bands[i] * std::pow(bands[i + 1] / bands[i], pos - i);
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
        for (double v : w) t.dequant_pool.push_back((float) (1.0 / v));
        if (rows != cols) symmetric = false;
        for (int a = 0; a < rows && symmetric; ++a)
          for (int b = 0; b < a; ++b)
            if ((float) (1.0 / w[(size_t) a * cols + b]) != (float) (1.0 / w[(size_t) b * cols + a])) {
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
