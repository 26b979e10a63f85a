// Rescale kernels (sm_100a): the two passes of the fixed-point separable resampler behind decodeSampled (resize.h).
// Vertical first, then horizontal, each output = saturate_u8((sum of u8 * Q15 weight + 2^14) >> 15) per channel with an
// 8-bit intermediate image -- exactly the arithmetic of the reference's weave_scale_u8 / pic-scale 0.7.6 path
// (/root/reference/jxlcoder/src/main/cpp/SizeScaler.cpp:38-144, /root/reference/weaver/src/scale.rs:294-361).
// Sources with alpha are premultiplied by the first pass that runs (on load) and divided back by the last one (on
// store); Nearest is a single gather.  Both passes are HBM-bound: the vertical pass reads each source row once per
// output row that covers it (the overlapping windows are served by L2) with fully coalesced uchar4 rows; the horizontal
// pass reads its window through L1.
#include <atomic>

#include "kernels.h"
#include "resize.h"

namespace jxlb {

extern std::atomic<uint64_t> g_launches_ac;

namespace {

__device__ __forceinline__ uint32_t SatU8(int32_t v) { return (uint32_t) (v < 0 ? 0 : v > 255 ? 255 : v); }

template <bool kPremulIn>
__device__ __forceinline__ void Accumulate(uchar4 v, int32_t wt, int32_t& r, int32_t& g, int32_t& b, int32_t& al) {
  if (kPremulIn) {
    r += wt * (int32_t) ResizePremul(v.x, v.w);
    g += wt * (int32_t) ResizePremul(v.y, v.w);
    b += wt * (int32_t) ResizePremul(v.z, v.w);
  } else {
    r += wt * v.x;
    g += wt * v.y;
    b += wt * v.z;
  }
  al += wt * v.w;
}

template <bool kUnpremulOut>
__device__ __forceinline__ uint32_t Finish(int32_t r, int32_t g, int32_t b, int32_t al) {
  uint32_t cr = SatU8(r >> 15), cg = SatU8(g >> 15), cb = SatU8(b >> 15);
  const uint32_t ca = SatU8(al >> 15);
  if (kUnpremulOut) {
    cr = ResizeUnpremul(cr, ca);
    cg = ResizeUnpremul(cg, ca);
    cb = ResizeUnpremul(cb, ca);
  }
  return cr | (cg << 8) | (cb << 16) | (ca << 24);
}

// thread = (pixel column x, output row y)
template <bool kPremulIn, bool kUnpremulOut>
__global__ void __launch_bounds__(256) ResizeVerticalKernel(const uint8_t* __restrict__ src, uint32_t src_stride, uint32_t width,
                                                            uint32_t out_h, ResizeAxisDev a, uint8_t* __restrict__ dst, uint32_t dst_stride) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= width || y >= out_h) return;
  const uint32_t start = a.start[y], n = a.count[y];
  const int16_t* w = a.weights + (size_t) y * a.taps;
  int32_t r = 1 << 14, g = 1 << 14, b = 1 << 14, al = 1 << 14;
  const uint8_t* p = src + (size_t) start * src_stride + 4 * (size_t) x;
  for (uint32_t t = 0; t < n; ++t)
    Accumulate<kPremulIn>(*reinterpret_cast<const uchar4*>(p + (size_t) t * src_stride), w[t], r, g, b, al);
  *reinterpret_cast<uint32_t*>(dst + (size_t) y * dst_stride + 4 * (size_t) x) = Finish<kUnpremulOut>(r, g, b, al);
}

// thread = (output column x, row y)
template <bool kPremulIn, bool kUnpremulOut>
__global__ void __launch_bounds__(256) ResizeHorizontalKernel(const uint8_t* __restrict__ src, uint32_t src_stride, uint32_t out_w,
                                                              uint32_t height, ResizeAxisDev a, uint8_t* __restrict__ dst,
                                                              uint32_t dst_stride) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= out_w || y >= height) return;
  const uint32_t start = a.start[x], n = a.count[x];
  const int16_t* w = a.weights + (size_t) x * a.taps;
  int32_t r = 1 << 14, g = 1 << 14, b = 1 << 14, al = 1 << 14;
  const uchar4* p = reinterpret_cast<const uchar4*>(src + (size_t) y * src_stride) + start;
  for (uint32_t t = 0; t < n; ++t) Accumulate<kPremulIn>(p[t], w[t], r, g, b, al);
  *reinterpret_cast<uint32_t*>(dst + (size_t) y * dst_stride + 4 * (size_t) x) = Finish<kUnpremulOut>(r, g, b, al);
}

__global__ void __launch_bounds__(256) ResizeNearestKernel(const uint8_t* __restrict__ src, uint32_t src_stride, uint32_t out_w, uint32_t out_h,
                                                           const uint32_t* __restrict__ row_of, const uint32_t* __restrict__ col_of,
                                                           uint8_t* __restrict__ dst, uint32_t dst_stride) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= out_w || y >= out_h) return;
  *reinterpret_cast<uint32_t*>(dst + (size_t) y * dst_stride + 4 * (size_t) x) =
      *reinterpret_cast<const uint32_t*>(src + (size_t) row_of[y] * src_stride + 4 * (size_t) col_of[x]);
}

}  // namespace

const uint8_t* LaunchResize(const ResizeDev& r, cudaStream_t stream) {
  const uint8_t* cur = r.src;
  uint32_t cur_stride = r.src_stride;
  if (r.nearest) {
    dim3 grid((r.scaled_w + 255) / 256, r.scaled_h, 1);
    ResizeNearestKernel<<<grid, 256, 0, stream>>>(cur, cur_stride, r.scaled_w, r.scaled_h, r.v.start, r.h.start, r.scaled, r.scaled_w * 4);
    ++g_launches_ac;
    return r.scaled;
  }
  if (r.has_v) {
    dim3 grid((r.src_w + 255) / 256, r.scaled_h, 1);
    const bool last = !r.has_h;
    auto k = !r.premultiply ? ResizeVerticalKernel<false, false> : last ? ResizeVerticalKernel<true, true> : ResizeVerticalKernel<true, false>;
    k<<<grid, 256, 0, stream>>>(cur, cur_stride, r.src_w, r.scaled_h, r.v, r.mid, r.src_w * 4);
    ++g_launches_ac;
    cur = r.mid;
    cur_stride = r.src_w * 4;
  }
  if (r.has_h) {
    dim3 grid((r.scaled_w + 255) / 256, r.scaled_h, 1);
    const bool first = !r.has_v;
    auto k = !r.premultiply ? ResizeHorizontalKernel<false, false> : first ? ResizeHorizontalKernel<true, true> : ResizeHorizontalKernel<false, true>;
    k<<<grid, 256, 0, stream>>>(cur, cur_stride, r.scaled_w, r.scaled_h, r.h, r.scaled, r.scaled_w * 4);
    ++g_launches_ac;
    cur = r.scaled;
  }
  return cur;  // with both passes skipped the source itself (stride r.src_stride) is the result
}

}  // namespace jxlb
