// Fused restoration-filter + colour + pack kernel (sm_100a): Gaborish -> EPF stage 0/1/2 -> XYB to output RGB ->
// transfer function -> dither -> alpha -> alpha association + Bitmap-format packing, one pass over HBM.
//
// A CTA owns a 64x16 output tile.  The XYB planes are read ONCE, with a halo of (1 if Gaborish) + 3/2/1 per EPF stage,
// into shared memory (mirrored at the image border exactly like the planar reference path in pixel_stages.h); the
// stages then ping-pong between two shared-memory buffers over shrinking regions, and the last stage's output feeds
// the colour conversion directly, so the 12 B/pixel intermediate planes between the stages never touch HBM.
// Algorithmic HBM traffic: 12 B/px read (+ halo) + 4 B/px written (RGBA_8888).
#include <atomic>
#include <cstdint>
#include <cstdlib>

#include "kernels.h"

namespace jxlb {

extern std::atomic<uint64_t> g_launches_ac;

namespace {

constexpr int kTW = 64, kTH = 16, kFilterThreads = 256;

// 3-channel tile in shared memory; (x, y) are tile-local coordinates.
struct SmemTile {
  const float* p;
  int stride, plane;
  __device__ float at(int c, int x, int y) const { return p[c * plane + y * stride + x]; }
};

struct EpfTileArgs {
  int gx0, gy0, W, H, off, rw, rh, stride, plane, cx_lo, cy_lo, ncx;
  const float* sig;
};

template <int kStage>
__device__ __forceinline__ void RunEpfStage(const SmemTile& in, float* out, const RestorationFilter& rf, const uint32_t* rcp11,
                                            const EpfTileArgs& a, int tid) {
  for (int i = tid; i < a.rw * a.rh; i += kFilterThreads) {
    const int lx = a.off + i % a.rw, ly = a.off + i / a.rw;
    const int mx = Mirror(a.gx0 + lx, a.W), my = Mirror(a.gy0 + ly, a.H);
    const float inv_sigma = a.sig[((my >> 3) - a.cy_lo) * a.ncx + (mx >> 3) - a.cx_lo];
    float o[3];
    EpfPixelT<kStage>(in, rf, rcp11, lx, ly, mx, my, inv_sigma, o);
    out[ly * a.stride + lx] = o[0];
    out[a.plane + ly * a.stride + lx] = o[1];
    out[2 * a.plane + ly * a.stride + lx] = o[2];
  }
}

struct FilterParams {
  PackParams pack;
  ColorParams cp;
  int32_t alpha_channel;
  uint32_t alpha_bits;
  uint32_t out16;
};

__global__ void __launch_bounds__(kFilterThreads) FilterColorKernel(const FrameDev f, const FilterParams fp, const NumericTables* nt,
                                                                    int halo) {
  if (*f.frame_bad) return;
  extern __shared__ __align__(16) float fsm[];
  const int bw = kTW + 2 * halo, bh = kTH + 2 * halo;
  const int stride = bw | 1;  // odd row stride
  const int plane = stride * bh;
  float* buf[2] = {fsm, fsm + 3 * plane};
  float* sig = fsm + 6 * plane;  // per-cell 1/sigma for the cells this tile can touch
  const int tx0 = blockIdx.x * kTW, ty0 = blockIdx.y * kTH;
  const int W = (int) f.width, H = (int) f.height;
  const int tid = threadIdx.x;
  // ---- per-cell inverse sigma for the clamped cell range of the tile
  const int cx_lo = (tx0 - halo < 0 ? 0 : tx0 - halo) >> 3, cy_lo = (ty0 - halo < 0 ? 0 : ty0 - halo) >> 3;
  const int cx_hi = ((tx0 + kTW + halo - 1 > W - 1 ? W - 1 : tx0 + kTW + halo - 1) >> 3);
  const int cy_hi = ((ty0 + kTH + halo - 1 > H - 1 ? H - 1 : ty0 + kTH + halo - 1) >> 3);
  const int ncx = cx_hi - cx_lo + 1, ncy = cy_hi - cy_lo + 1;
  if (f.rf.epf_iters)
    for (int i = tid; i < ncx * ncy; i += kFilterThreads) {
      const size_t ci = (size_t) (cy_lo + i / ncx) * f.w8 + cx_lo + i % ncx;
      sig[i] = EpfInvSigma(f, f.cell_hfmul[ci], f.cell_sharp[ci]);
    }
  // ---- load the tile + halo (mirrored at the image border)
  const size_t pplane = (size_t) f.plane_h * f.plane_stride;
  for (int i = tid; i < bw * bh; i += kFilterThreads) {
    const int lx = i % bw, ly = i / bw;
    const int gx = Mirror(tx0 - halo + lx, W), gy = Mirror(ty0 - halo + ly, H);
    const size_t go = (size_t) gy * f.plane_stride + gx;
    buf[0][ly * stride + lx] = f.xyb0[go];
    buf[0][plane + ly * stride + lx] = f.xyb0[pplane + go];
    buf[0][2 * plane + ly * stride + lx] = f.xyb0[2 * pplane + go];
  }
  __syncthreads();
  int cur = 0, off = 0;  // `off`: how much of the halo has been consumed
  // ---- stages
  for (int stage = -1; stage < 3; ++stage) {
    int radius;
    if (stage < 0) {
      if (!f.rf.gab) continue;
      radius = 1;
    } else {
      const int iters = f.rf.epf_iters;
      const bool run = (stage == 0 && iters == 3) || (stage == 1 && iters >= 1) || (stage == 2 && iters >= 2);
      if (!run) continue;
      radius = stage == 0 ? 3 : stage == 1 ? 2 : 1;
    }
    off += radius;
    const int rw = bw - 2 * off, rh = bh - 2 * off;
    SmemTile in{buf[cur], stride, plane};
    float* out = buf[cur ^ 1];
    if (stage < 0) {
      for (int i = tid; i < rw * rh; i += kFilterThreads) {
        const int lx = off + i % rw, ly = off + i / rw;
#pragma unroll
        for (int c = 0; c < 3; ++c) out[c * plane + ly * stride + lx] = GaborishSample(in, c, lx, ly, f.rf.gab_w1[c], f.rf.gab_w2[c]);
      }
    } else {
      const EpfTileArgs ea{tx0 - halo, ty0 - halo, W, H, off, rw, rh, stride, plane, cx_lo, cy_lo, ncx, sig};
      if (stage == 0) RunEpfStage<0>(in, out, f.rf, nt->rcp11, ea, tid);
      else if (stage == 1) RunEpfStage<1>(in, out, f.rf, nt->rcp11, ea, tid);
      else RunEpfStage<2>(in, out, f.rf, nt->rcp11, ea, tid);
    }
    cur ^= 1;
    __syncthreads();
  }
  // ---- colour + pack
  const float* fin = buf[cur];
  for (int i = tid; i < kTW * kTH; i += kFilterThreads) {
    const int x = tx0 + i % kTW, y = ty0 + i / kTW;
    if (x >= W || y >= H) continue;
    const int lx = halo + i % kTW, ly = halo + i / kTW;
    float rgb[3];
    XybToEncodedRgb(fin[ly * stride + lx], fin[plane + ly * stride + lx], fin[2 * plane + ly * stride + lx], fp.cp, rgb);
    uint32_t v[3];
    const uint32_t maxout = fp.out16 ? 65535u : 255u;
    if (fp.out16) {
      for (int c = 0; c < 3; ++c) {
        float s = rgb[c] * 65535.0f;
        s = s < 0.0f ? 0.0f : s > 65535.0f ? 65535.0f : s;
        v[c] = (uint32_t) rintf(s);
      }
    } else {
      const float d = nt->dither[DitherIndex(f, (uint32_t) x, (uint32_t) y)];
      for (int c = 0; c < 3; ++c) v[c] = ToU8Dithered(rgb[c], d);
    }
    if (fp.cp.grey) v[0] = v[2] = v[1];
    uint32_t a = maxout;
    if (fp.alpha_channel >= 0)
      a = ScaleSample(f.mod[(size_t) fp.alpha_channel * f.height * f.mod_stride + (size_t) y * f.mod_stride + x], fp.alpha_bits, maxout);
    PackRgba(fp.pack, (uint32_t) x, (uint32_t) y, v[0], v[1], v[2], a);
  }
}


// ---- fast path: Gaborish + one EPF iteration (stage 1) + colour + pack ------------------------------------------------
// The configuration libjxl's encoder emits at distance ~1 (gab = 1, epf_iters = 1).  Same arithmetic, operation by
// operation, as the generic kernel above (GaborishSample / EpfPixelT<1> / XybToEncodedRgb), restructured around
// 4-pixel horizontal strips held in registers: every thread produces 4 adjacent outputs per stage from 128-bit shared
// memory loads with compile-time offsets, instead of one output from ~25 scalar loads per channel with run-time strides.
//   tile: 64 x 16 outputs; Gaborish output 68 x 20; input 70 x 22.  Shared rows are 84 floats: index = column + 8.
constexpr int kSW = 84, kSH = 22, kSPlane = kSW * kSH;
constexpr int kFastSmemFloats = 6 * kSPlane + 16;

__device__ __forceinline__ float4 Lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float2 Lds2(const float* p) { return *reinterpret_cast<const float2*>(p); }

__global__ void __launch_bounds__(kFilterThreads, 3) FilterColorFastKernel(const FrameDev f, const FilterParams fp, const NumericTables* nt) {
  if (*f.frame_bad) return;
  extern __shared__ __align__(16) float fsm[];
  float* in0 = fsm;                  // [3][kSH][kSW] input XYB (rows ty0-3 .., cols tx0-8 ..)
  float* gb = fsm + 3 * kSPlane;     // Gaborish output, same geometry
  float* sig = fsm + 6 * kSPlane;    // 1/sigma of the tile's 8 x 2 cells
  const int tx0 = blockIdx.x * kTW, ty0 = blockIdx.y * kTH;
  const int W = (int) f.width, H = (int) f.height;
  const int tid = threadIdx.x;
  if (tid < 16) {
    const int cx = (tx0 >> 3) + (tid & 7), cy = (ty0 >> 3) + (tid >> 3);
    float v = 0.0f;
    if (cx < (int) f.w8 && cy < (int) f.h8) {
      const size_t ci = (size_t) cy * f.w8 + cx;
      v = EpfInvSigma(f, f.cell_hfmul[ci], f.cell_sharp[ci]);
    }
    sig[tid] = v;
  }
  // ---- load input: rows ty0-3 .. ty0+18, columns tx0-4 .. tx0+67 (needed: -3 .. 66)
  const size_t pplane = (size_t) f.plane_h * f.plane_stride;
  const bool interior = tx0 >= 4 && tx0 + 68 <= W && ty0 >= 3 && ty0 + 19 <= H;
  if (interior) {
    // 3 x 22 x 18 float4 = 1188 loads over 256 threads: all five of a thread issued before the first store
    constexpr int kLoads = 3 * kSH * 18, kPer = (kLoads + kFilterThreads - 1) / kFilterThreads;
    float4 v[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int i = tid + k * kFilterThreads;
      if (i < kLoads) {
        const int q = i % 18, rc = i / 18, r = rc % kSH, c = rc / kSH;
        v[k] = __ldg(reinterpret_cast<const float4*>(f.xyb0 + c * pplane + (size_t) (ty0 - 3 + r) * f.plane_stride + tx0 - 4 + 4 * q));
      }
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int i = tid + k * kFilterThreads;
      if (i < kLoads) {
        const int q = i % 18, rc = i / 18, r = rc % kSH, c = rc / kSH;
        *reinterpret_cast<float4*>(in0 + c * kSPlane + r * kSW + 4 + 4 * q) = v[k];
      }
    }
  } else {
    for (int i = tid; i < kSH * 72; i += kFilterThreads) {
      const int lx = i % 72, r = i / 72;
      const int gx = Mirror(tx0 - 4 + lx, W), gy = Mirror(ty0 - 3 + r, H);
      const size_t go = (size_t) gy * f.plane_stride + gx;
      in0[r * kSW + 4 + lx] = f.xyb0[go];
      in0[kSPlane + r * kSW + 4 + lx] = f.xyb0[pplane + go];
      in0[2 * kSPlane + r * kSW + 4 + lx] = f.xyb0[2 * pplane + go];
    }
  }
  __syncthreads();
  // ---- Gaborish: 18 strips (columns -4 .. 67) x 20 rows (-2 .. 17); shared row index = row + 3
  for (int i = tid; i < 18 * 20; i += kFilterThreads) {
    const int strip = i % 18, r = 1 + i / 18;  // r: shared row of the output
    const int sx = 4 + 4 * strip;              // shared column of the strip's first output
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float w1 = f.rf.gab_w1[c], w2 = f.rf.gab_w2[c];
      const float norm = 1.0f / (1.0f + 4.0f * (w1 + w2));
      const float* p = in0 + c * kSPlane + r * kSW + sx;
      float up[6], mid[6], dn[6];
      {
        const float4 a = Lds4(p - kSW), b = Lds4(p), d = Lds4(p + kSW);
        up[0] = p[-kSW - 1]; up[1] = a.x; up[2] = a.y; up[3] = a.z; up[4] = a.w; up[5] = p[-kSW + 4];
        mid[0] = p[-1]; mid[1] = b.x; mid[2] = b.y; mid[3] = b.z; mid[4] = b.w; mid[5] = p[4];
        dn[0] = p[kSW - 1]; dn[1] = d.x; dn[2] = d.y; dn[3] = d.z; dn[4] = d.w; dn[5] = p[kSW + 4];
      }
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float centre = mid[1 + j];
        const float cross = up[1 + j] + dn[1 + j] + mid[j] + mid[2 + j];
        const float diag = up[j] + up[2 + j] + dn[j] + dn[2 + j];
        o[j] = (centre + w1 * cross + w2 * diag) * norm;
      }
      *reinterpret_cast<float4*>(gb + c * kSPlane + r * kSW + sx) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  __syncthreads();
  // ---- EPF stage 1 on a 4-pixel strip per thread
  const int strip = tid & 15, ry = tid >> 4;      // output row ry (0 .. 15), columns 4 * strip .. + 3
  const int sx = 8 + 4 * strip, sr = ry + 3;      // shared column / row of the strip
  const int x0 = tx0 + 4 * strip, y = ty0 + ry;
  const float inv_sigma = sig[(ry >> 3) * 8 + (strip >> 1)];
  float outv[3][4];
  if (inv_sigma < kEpfSkipThreshold) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 v = Lds4(gb + c * kSPlane + sr * kSW + sx);
      outv[c][0] = v.x; outv[c][1] = v.y; outv[c][2] = v.z; outv[c][3] = v.w;
    }
  } else {
    float sad[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) sad[j][i] = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float sc = f.rf.epf_channel_scale[c];
      const float* p = gb + c * kSPlane + sr * kSW + sx;
      float r0[8], rm1[6], rp1[6], rm2[4], rp2[4];
      {
        const float2 a = Lds2(p - 2), e = Lds2(p + 4);
        const float4 b = Lds4(p);
        r0[0] = a.x; r0[1] = a.y; r0[2] = b.x; r0[3] = b.y; r0[4] = b.z; r0[5] = b.w; r0[6] = e.x; r0[7] = e.y;
        const float4 m = Lds4(p - kSW), q = Lds4(p + kSW);
        rm1[0] = p[-kSW - 1]; rm1[1] = m.x; rm1[2] = m.y; rm1[3] = m.z; rm1[4] = m.w; rm1[5] = p[-kSW + 4];
        rp1[0] = p[kSW - 1]; rp1[1] = q.x; rp1[2] = q.y; rp1[3] = q.z; rp1[4] = q.w; rp1[5] = p[kSW + 4];
        const float4 m2 = Lds4(p - 2 * kSW), q2 = Lds4(p + 2 * kSW);
        rm2[0] = m2.x; rm2[1] = m2.y; rm2[2] = m2.z; rm2[3] = m2.w;
        rp2[0] = q2.x; rp2[1] = q2.y; rp2[2] = q2.z; rp2[3] = q2.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // centre plus-shape: cc0 = r0[2+j], up = rm1[1+j], down = rp1[1+j], left = r0[1+j], right = r0[3+j]
        const float cc0 = r0[2 + j], ccu = rm1[1 + j], ccd = rp1[1 + j], ccl = r0[1 + j], ccr = r0[3 + j];
        float s;
        // neighbour (0, -1)
        s = fabsf(rm1[1 + j] - cc0);
        s += fabsf(rm2[j] - ccu);
        s += fabsf(r0[2 + j] - ccd);
        s += fabsf(rm1[j] - ccl);
        s += fabsf(rm1[2 + j] - ccr);
        sad[j][0] += s * sc;
        // neighbour (0, 1)
        s = fabsf(rp1[1 + j] - cc0);
        s += fabsf(r0[2 + j] - ccu);
        s += fabsf(rp2[j] - ccd);
        s += fabsf(rp1[j] - ccl);
        s += fabsf(rp1[2 + j] - ccr);
        sad[j][1] += s * sc;
        // neighbour (-1, 0)
        s = fabsf(r0[1 + j] - cc0);
        s += fabsf(rm1[j] - ccu);
        s += fabsf(rp1[j] - ccd);
        s += fabsf(r0[j] - ccl);
        s += fabsf(r0[2 + j] - ccr);
        sad[j][2] += s * sc;
        // neighbour (1, 0)
        s = fabsf(r0[3 + j] - cc0);
        s += fabsf(rm1[2 + j] - ccu);
        s += fabsf(rp1[2 + j] - ccd);
        s += fabsf(r0[2 + j] - ccl);
        s += fabsf(r0[4 + j] - ccr);
        sad[j][3] += s * sc;
      }
    }
    float wgt[4][4], inv[4];
    const bool yborder = (y & 7) == 0 || (y & 7) == 7;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float sm = 1.65f;
      const int xm = (x0 + j) & 7;
      if (xm == 0 || xm == 7 || yborder) sm *= f.rf.epf_border_sad_mul;
      const float isg = inv_sigma * sm;
      float wsum = 1.0f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float w = 1.0f + sad[j][i] * isg;
        if (w < 0.0f) w = 0.0f;
        wsum += w;
        wgt[j][i] = w;
      }
      inv[j] = ApproxRcp(nt->rcp11, wsum);  // libjxl's ApproximateReciprocal (JXL_HIGH_PRECISION=0)
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* p = gb + c * kSPlane + sr * kSW + sx;
      const float4 m = Lds4(p - kSW), b = Lds4(p), q = Lds4(p + kSW);
      const float up[4] = {m.x, m.y, m.z, m.w}, dn[4] = {q.x, q.y, q.z, q.w};
      const float mid[6] = {p[-1], b.x, b.y, b.z, b.w, p[4]};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = mid[1 + j];
        a += wgt[j][0] * up[j];
        a += wgt[j][1] * dn[j];
        a += wgt[j][2] * mid[j];
        a += wgt[j][3] * mid[2 + j];
        outv[c][j] = a * inv[j];
      }
    }
  }
  // ---- colour + pack
  if (y >= H || x0 >= W) return;
  uint32_t px[4][4];
  float dith[4] = {0.f, 0.f, 0.f, 0.f};
  if (!fp.out16) {
    if (f.orientation == 1 && (f.dither_x0 | f.dither_y0) == 0) {
      const float4 d = *reinterpret_cast<const float4*>(nt->dither + (y & 31) * 32 + (x0 & 31));
      dith[0] = d.x; dith[1] = d.y; dith[2] = d.z; dith[3] = d.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) dith[j] = nt->dither[DitherIndex(f, (uint32_t) (x0 + j), (uint32_t) y)];
    }
  }
  const uint32_t maxout = fp.out16 ? 65535u : 255u;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float rgb[3];
    XybToEncodedRgb(outv[0][j], outv[1][j], outv[2][j], fp.cp, rgb);
    if (fp.out16) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float sv = rgb[c] * 65535.0f;
        sv = sv < 0.0f ? 0.0f : sv > 65535.0f ? 65535.0f : sv;
        px[j][c] = (uint32_t) rintf(sv);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) px[j][c] = ToU8Dithered(rgb[c], dith[j]);
    }
    if (fp.cp.grey) px[j][0] = px[j][2] = px[j][1];
    px[j][3] = maxout;
  }
  if (fp.alpha_channel >= 0) {
    const int32_t* arow = f.mod + (size_t) fp.alpha_channel * f.height * f.mod_stride + (size_t) y * f.mod_stride;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (x0 + j < W) px[j][3] = ScaleSample(arow[x0 + j], fp.alpha_bits, maxout);
  }
  if (!fp.out16 && (fp.pack.format == 0 || fp.pack.format == 4) && !fp.pack.associate && x0 + 3 < W) {
    uint4 o;
    o.x = px[0][0] | (px[0][1] << 8) | (px[0][2] << 16) | (px[0][3] << 24);
    o.y = px[1][0] | (px[1][1] << 8) | (px[1][2] << 16) | (px[1][3] << 24);
    o.z = px[2][0] | (px[2][1] << 8) | (px[2][2] << 16) | (px[2][3] << 24);
    o.w = px[3][0] | (px[3][1] << 8) | (px[3][2] << 16) | (px[3][3] << 24);
    *reinterpret_cast<uint4*>(fp.pack.dst + (size_t) y * fp.pack.dst_stride + 4 * (size_t) x0) = o;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (x0 + j < W) PackRgba(fp.pack, (uint32_t) (x0 + j), (uint32_t) y, px[j][0], px[j][1], px[j][2], px[j][3]);
  }
}

}  // namespace

void LaunchFilterColorPack(const FrameDev& f, const ColorParams& cp, const NumericTables* nt_dev, const OutputDesc& od,
                           const PackParams& pack, cudaStream_t stream) {
  int halo = f.rf.gab ? 1 : 0;
  if (f.rf.epf_iters == 3) halo += 3;
  if (f.rf.epf_iters >= 1) halo += 2;
  if (f.rf.epf_iters >= 2) halo += 1;
  const int bw = kTW + 2 * halo, bh = kTH + 2 * halo;
  const int stride = bw | 1;
  const int ncells = ((kTW + 2 * halo) / 8 + 2) * ((kTH + 2 * halo) / 8 + 2);
  const size_t smem = ((size_t) 6 * stride * bh + ncells) * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(FilterColorKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    // same (maximal) shared-memory carve-out for every long-running kernel: CTAs of kernels with different carve-outs
    // cannot share an SM, which would serialise the batches that overlap on different streams
    cudaFuncSetAttribute(FilterColorKernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = smem;
  }
  FilterParams fp;
  fp.pack = pack;
  fp.cp = cp;
  fp.alpha_channel = od.alpha_channel;
  fp.alpha_bits = od.alpha_bits;
  fp.out16 = od.bits16;
  dim3 grid((f.width + kTW - 1) / kTW, (f.height + kTH - 1) / kTH, 1);
  static const bool no_fast = getenv("JXLB_NO_FAST_FILTER") != nullptr;
  // the uint4 store of the fast path needs 16-byte aligned output rows
  const bool aligned = (reinterpret_cast<uintptr_t>(pack.dst) & 15) == 0 && (pack.dst_stride & 15) == 0;
  if (f.rf.gab && f.rf.epf_iters == 1 && aligned && !no_fast) {
    static bool fast_configured = false;
    if (!fast_configured) {
      cudaFuncSetAttribute(FilterColorFastKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (kFastSmemFloats * sizeof(float)));
      cudaFuncSetAttribute(FilterColorFastKernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      fast_configured = true;
    }
    FilterColorFastKernel<<<grid, kFilterThreads, kFastSmemFloats * sizeof(float), stream>>>(f, fp, nt_dev);
    ++g_launches_ac;
    return;
  }
  FilterColorKernel<<<grid, kFilterThreads, smem, stream>>>(f, fp, nt_dev, halo);
  ++g_launches_ac;
}

}  // namespace jxlb
