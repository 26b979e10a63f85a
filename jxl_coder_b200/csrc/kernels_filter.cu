// Fused restoration-filter + colour + pack kernel (sm_100a): Gaborish -> EPF stage 0/1/2 -> XYB to output RGB ->
// transfer function -> dither -> alpha -> alpha association + Bitmap-format packing, one pass over HBM.
//
// A CTA owns a 64x16 output tile.  The XYB planes are read ONCE, with a halo of (1 if Gaborish) + 3/2/1 per EPF stage,
// into shared memory (mirrored at the image border exactly like the planar reference path in pixel_stages.h); the
// stages then ping-pong between two shared-memory buffers over shrinking regions, and the last stage's output feeds
// the colour conversion directly, so the 12 B/pixel intermediate planes between the stages never touch HBM.
// Algorithmic HBM traffic: 12 B/px read (+ halo) + 4 B/px written (RGBA_8888).
#include <atomic>

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
__device__ __forceinline__ void RunEpfStage(const SmemTile& in, float* out, const RestorationFilter& rf, const EpfTileArgs& a, int tid) {
  for (int i = tid; i < a.rw * a.rh; i += kFilterThreads) {
    const int lx = a.off + i % a.rw, ly = a.off + i / a.rw;
    const int mx = Mirror(a.gx0 + lx, a.W), my = Mirror(a.gy0 + ly, a.H);
    const float inv_sigma = a.sig[((my >> 3) - a.cy_lo) * a.ncx + (mx >> 3) - a.cx_lo];
    float o[3];
    EpfPixelT<kStage>(in, rf, lx, ly, mx, my, inv_sigma, o);
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
      if (stage == 0) RunEpfStage<0>(in, out, f.rf, ea, tid);
      else if (stage == 1) RunEpfStage<1>(in, out, f.rf, ea, tid);
      else RunEpfStage<2>(in, out, f.rf, ea, tid);
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
      const float d = nt->dither[(y & 31) * 32 + (x & 31)];
      for (int c = 0; c < 3; ++c) v[c] = ToU8Dithered(rgb[c], d);
    }
    if (fp.cp.grey) v[0] = v[2] = v[1];
    uint32_t a = maxout;
    if (fp.alpha_channel >= 0)
      a = ScaleSample(f.mod[(size_t) fp.alpha_channel * f.height * f.mod_stride + (size_t) y * f.mod_stride + x], fp.alpha_bits, maxout);
    PackRgba(fp.pack, (uint32_t) x, (uint32_t) y, v[0], v[1], v[2], a);
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
  FilterColorKernel<<<grid, kFilterThreads, smem, stream>>>(f, fp, nt_dev, halo);
  ++g_launches_ac;
}

}  // namespace jxlb
