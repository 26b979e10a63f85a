// Kernels around the main decode chain (sm_100a), all simple and HBM- or latency-bound:
//   * ColorMatrixKernel   -- the api_level < 34 colour pass (color_matrix.h): in place on straight RGBA8, one thread per
//                            pixel, both LUTs (1 KB + 2 KB) staged in shared memory; 4 B read + 4 B written per pixel;
//   * OrientKernel        -- codestream orientation 2..8 on the decoded picture;
//   * PlaceKernel         -- a cropped frame laid over a cleared canvas;
//   * Squeeze*Kernel      -- inverse squeeze of lossy extra channels (squeeze.h), one launch per step.
#include <atomic>

#include "color_matrix.h"
#include "kernels.h"
#include "squeeze.h"

namespace jxlb {

extern std::atomic<uint64_t> g_launches_ac;

namespace {

// Rec2408ToneMapper::transferTone stops tone-mapping a row at its first pixel of zero luminance (color_matrix.h): this
// kernel finds that pixel for every row (width when there is none).  kBits16: RGBA16 source with the 2^16-entry table.
template <bool kBits16>
__global__ void __launch_bounds__(256) FirstBlackKernel(const uint8_t* __restrict__ img, uint32_t stride, uint32_t width, const float* __restrict__ lin,
                                                        uint32_t* __restrict__ row_first) {
  const uint32_t y = blockIdx.x;
  __shared__ uint32_t first;
  if (threadIdx.x == 0) first = width;
  __syncthreads();
  uint32_t mine = width;
  for (uint32_t x = threadIdx.x; x < width && x < mine; x += blockDim.x) {
    float r, g, b;
    if (kBits16) {
      const uint2 v = *reinterpret_cast<const uint2*>(img + (size_t) y * stride + (size_t) x * 8);
      r = lin[v.x & 0xFFFF]; g = lin[v.x >> 16]; b = lin[v.y & 0xFFFF];
    } else {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(img + (size_t) y * stride + (size_t) x * 4);
      r = lin[v & 0xFF]; g = lin[(v >> 8) & 0xFF]; b = lin[(v >> 16) & 0xFF];
    }
    const float light = __fadd_rn(__fadd_rn(__fmul_rn(0.2627f, r), __fmul_rn(0.6780f, g)), __fmul_rn(0.0593f, b));
    if (light == 0.0f) mine = x;
  }
  if (mine < width) atomicMin(&first, mine);
  __syncthreads();
  if (threadIdx.x == 0) row_first[y] = first;
}

// One pixel of the pass after the linearisation: [tone map] -> matrix; separate multiplies and adds, as the reference's
// scalar loops (no FMA contraction).
__device__ __forceinline__ void ToneMapAndMatrix(float r, float g, float b, bool tone, float wa, float wb, const float* m, float out[3]) {
  if (tone) {
    const float light = __fadd_rn(__fadd_rn(__fmul_rn(0.2627f, r), __fmul_rn(0.6780f, g)), __fmul_rn(0.0593f, b));
    const float scale = __fdiv_rn(__fadd_rn(1.f, __fmul_rn(wa, light)), __fadd_rn(1.f, __fmul_rn(wb, light)));
    r = fminf(__fmul_rn(r, scale), 1.f);
    g = fminf(__fmul_rn(g, scale), 1.f);
    b = fminf(__fmul_rn(b, scale), 1.f);
  }
  out[0] = __fadd_rn(__fadd_rn(__fmul_rn(r, m[0]), __fmul_rn(g, m[1])), __fmul_rn(b, m[2]));
  out[1] = __fadd_rn(__fadd_rn(__fmul_rn(r, m[3]), __fmul_rn(g, m[4])), __fmul_rn(b, m[5]));
  out[2] = __fadd_rn(__fadd_rn(__fmul_rn(r, m[6]), __fmul_rn(g, m[7])), __fmul_rn(b, m[8]));
}

__global__ void __launch_bounds__(256) ColorMatrixKernel(uint8_t* __restrict__ img, uint32_t stride, uint32_t width, uint32_t height,
                                                         const ColorMatrixPlan* __restrict__ plan, const uint32_t* __restrict__ row_first) {
  __shared__ float lin[256];
  __shared__ uint8_t gam[2052];
  __shared__ float m[9];
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) lin[i] = plan->linearize[i];
  for (uint32_t i = threadIdx.x; i < 2049; i += blockDim.x) gam[i] = plan->gamma[i];
  if (threadIdx.x < 9) m[threadIdx.x] = plan->m[threadIdx.x];
  __syncthreads();
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= width) return;
  const bool tonemap = plan->tonemap != 0;
  const float wa = plan->weight_a, wb = plan->weight_b;
  for (uint32_t y = blockIdx.y; y < height; y += gridDim.y) {
    uint32_t* px = reinterpret_cast<uint32_t*>(img + (size_t) y * stride) + x;
    const uint32_t v = *px;
    float o[3];
    ToneMapAndMatrix(lin[v & 0xFF], lin[(v >> 8) & 0xFF], lin[(v >> 16) & 0xFF], tonemap && x < row_first[y], wa, wb, m, o);
    const uint32_t ir = min((uint32_t) (fminf(fmaxf(o[0], 0.f), 1.f) * 2048.f), 2048u);
    const uint32_t ig = min((uint32_t) (fminf(fmaxf(o[1], 0.f), 1.f) * 2048.f), 2048u);
    const uint32_t ib = min((uint32_t) (fminf(fmaxf(o[2], 0.f), 1.f) * 2048.f), 2048u);
    *px = (uint32_t) gam[ir] | ((uint32_t) gam[ig] << 8) | ((uint32_t) gam[ib] << 16) | (v & 0xFF000000u);
  }
}

// applyColorMatrix16Bit: RGBA16 in place, 2^16-entry tables read through L1 / L2 (384 KB per image).
__global__ void __launch_bounds__(256) ColorMatrix16Kernel(uint8_t* __restrict__ img, uint32_t stride, uint32_t width, uint32_t height,
                                                           const ColorMatrixPlan* __restrict__ plan, const ColorMatrixTables16* __restrict__ t,
                                                           const uint32_t* __restrict__ row_first) {
  __shared__ float m[9];
  if (threadIdx.x < 9) m[threadIdx.x] = plan->m[threadIdx.x];
  __syncthreads();
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= width) return;
  const bool tonemap = plan->tonemap != 0;
  const float wa = plan->weight_a, wb = plan->weight_b;
  for (uint32_t y = blockIdx.y; y < height; y += gridDim.y) {
    uint2* px = reinterpret_cast<uint2*>(img + (size_t) y * stride) + x;
    const uint2 v = *px;
    float o[3];
    ToneMapAndMatrix(__ldg(t->linearize + (v.x & 0xFFFF)), __ldg(t->linearize + (v.x >> 16)), __ldg(t->linearize + (v.y & 0xFFFF)),
                     tonemap && x < row_first[y], wa, wb, m, o);
    const uint32_t ir = min((uint32_t) (fminf(fmaxf(o[0], 0.f), 1.f) * 65535.f), 65535u);
    const uint32_t ig = min((uint32_t) (fminf(fmaxf(o[1], 0.f), 1.f) * 65535.f), 65535u);
    const uint32_t ib = min((uint32_t) (fminf(fmaxf(o[2], 0.f), 1.f) * 65535.f), 65535u);
    uint2 w;
    w.x = (uint32_t) __ldg(t->gamma + ir) | ((uint32_t) __ldg(t->gamma + ig) << 16);
    w.y = (uint32_t) __ldg(t->gamma + ib) | (v.y & 0xFFFF0000u);
    *px = w;
  }
}

// Orientation: thread per OUTPUT pixel (coalesced stores; the transposing cases read with a stride, served by L2 -- the
// pass only runs for files that carry a non-identity orientation).
template <typename Pixel>
__global__ void __launch_bounds__(256) OrientKernel(const uint8_t* __restrict__ src, uint32_t src_stride, uint32_t w, uint32_t h,
                                                    uint32_t orientation, uint8_t* __restrict__ dst, uint32_t dst_stride) {
  const uint32_t ow = orientation >= 5 ? h : w, oh = orientation >= 5 ? w : h;
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= ow || y >= oh) return;
  uint32_t sx, sy;
  switch (orientation) {
    case 2: sx = w - 1 - x; sy = y; break;          // flip horizontal
    case 3: sx = w - 1 - x; sy = h - 1 - y; break;  // rotate 180
    case 4: sx = x; sy = h - 1 - y; break;          // flip vertical
    case 5: sx = y; sy = x; break;                  // transpose
    case 6: sx = y; sy = h - 1 - x; break;          // rotate 90 clockwise
    case 7: sx = w - 1 - y; sy = h - 1 - x; break;  // anti-transpose
    case 8: sx = w - 1 - y; sy = x; break;          // rotate 90 counter-clockwise
    default: sx = x; sy = y; break;
  }
  reinterpret_cast<Pixel*>(dst + (size_t) y * dst_stride)[x] = reinterpret_cast<const Pixel*>(src + (size_t) sy * src_stride)[sx];
}

// ---- inverse squeeze of lossy extra channels (squeeze.h) ----
// The channels of the global stream were decoded on the host (they are tiny) and arrive in the const region: one CTA
// per channel copies them to their place in the squeeze buffer.
__global__ void __launch_bounds__(256) SqueezeScatterGlobalKernel(const FrameDev f) {
  if (*f.frame_bad) return;
  const uint32_t c = blockIdx.x;
  if (c >= f.sq_global) return;
  size_t o = 0;
  for (uint32_t k = 0; k < c; ++k) o += (size_t) f.sq_ch[k].w * f.sq_ch[k].h;
  const SqChannel sc = f.sq_ch[c];
  const size_t n = (size_t) sc.w * sc.h;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) f.sq_buf[sc.off + i] = f.sq_global_data[o + i];
}

// One inverse step: a thread per row (horizontal step: serial along x because the tendency term looks at the pixel
// just reconstructed) or per column (vertical step, coalesced).  A 4096^2 alpha plane takes ~20 steps of at most
// 4096 threads each: latency-bound, a few hundred microseconds in total, off the colour path.
__global__ void __launch_bounds__(128) SqueezeStepKernel(const FrameDev f, uint32_t k) {
  if (*f.frame_bad) return;
  const SqStep st = f.sq_steps[k];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool final_plane = st.out_off == 0xFFFFFFFFu;
  int32_t* out = final_plane ? f.mod + (size_t) st.final_channel * f.height * f.mod_stride : f.sq_buf + st.out_off;
  const uint32_t ow = st.horizontal ? st.avg_w + st.res_w : st.avg_w;
  const uint32_t ostride = final_plane ? f.mod_stride : ow;
  if (st.horizontal) {
    if (i >= st.avg_h) return;
    InvSqueezeRow(f.sq_buf + st.avg_off + (size_t) i * st.avg_w, f.sq_buf + st.res_off + (size_t) i * st.res_w, out + (size_t) i * ostride,
                  st.avg_w, st.res_w);
  } else {
    if (i >= st.avg_w) return;
    InvSqueezeColumn(f.sq_buf + st.avg_off + i, st.avg_w, f.sq_buf + st.res_off + i, st.res_w, out + i, ostride, st.avg_h, st.res_h);
  }
}

// A cropped frame laid over a cleared canvas: thread per canvas pixel.
template <typename Pixel>
__global__ void __launch_bounds__(256) PlaceKernel(const uint8_t* __restrict__ src, uint32_t src_stride, uint32_t fw, uint32_t fh, int32_t x0,
                                                   int32_t y0, Pixel fill, uint8_t* __restrict__ dst, uint32_t dst_stride, uint32_t cw, uint32_t ch) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= cw || y >= ch) return;
  const int64_t fx = (int64_t) x - x0, fy = (int64_t) y - y0;
  Pixel v = fill;
  if (fx >= 0 && fy >= 0 && fx < (int64_t) fw && fy < (int64_t) fh) v = reinterpret_cast<const Pixel*>(src + (size_t) fy * src_stride)[fx];
  reinterpret_cast<Pixel*>(dst + (size_t) y * dst_stride)[x] = v;
}

}  // namespace

namespace {
// Frame composition (JPEG XL blending, ISO/IEC 18181-1 F.? / libjxl blending.cc): one thread per canvas pixel.  Samples are
// straight RGBA8 / RGBA16; the arithmetic is float on v * (1 / max) like libjxl's, rounded back with lrintf.
template <typename Pixel, int kMax>
__global__ void __launch_bounds__(256) CompositeKernel(const CompositeParams p) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= p.cw || y >= p.ch) return;
  // reference slots hold unrounded float samples in [0, 1] like libjxl's (an animation blends many frames in a row)
  float4 bgf = make_float4(0.f, 0.f, 0.f, p.has_alpha ? 0.f : 1.f);
  if (p.bg) bgf = p.bg[(size_t) y * p.cw + x];
  float of[4] = {bgf.x, bgf.y, bgf.z, bgf.w};
  bool blended = false;
  const int32_t fx = (int32_t) x - p.x0, fy = (int32_t) y - p.y0;
  if (fx >= 0 && fy >= 0 && fx < (int32_t) p.fw && fy < (int32_t) p.fh) {
    uint32_t fg[4];
    const Pixel v = *reinterpret_cast<const Pixel*>(p.fg + (size_t) fy * p.fg_stride + (size_t) fx * sizeof(Pixel));
    if (sizeof(Pixel) == 4) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(&v);
      fg[0] = w & 0xFF; fg[1] = (w >> 8) & 0xFF; fg[2] = (w >> 16) & 0xFF; fg[3] = w >> 24;
    } else {
      const uint2 w = *reinterpret_cast<const uint2*>(&v);
      fg[0] = w.x & 0xFFFF; fg[1] = w.x >> 16; fg[2] = w.y & 0xFFFF; fg[3] = w.y >> 16;
    }
    blended = true;
    {
      // separately rounded operations in libjxl's order (its blending loops are scalar code without FMA)
      const float inv = 1.0f / (float) kMax;
      float fa = p.has_alpha ? __fmul_rn((float) fg[3], inv) : 1.0f;
      const float ba = bgf.w;
      if (p.clamp) fa = fminf(fmaxf(fa, 0.0f), 1.0f);
      const float one_m_fa = __fsub_rn(1.0f, fa);
      const float na = __fsub_rn(1.0f, __fmul_rn(one_m_fa, __fsub_rn(1.0f, ba)));
      const float rna = na > 0.0f ? __fdiv_rn(1.0f, na) : 0.0f;
      const float bgc[3] = {bgf.x, bgf.y, bgf.z};
      float oc[3];
      for (int c = 0; c < 3; ++c) {
        const float f = __fmul_rn((float) fg[c], inv), b = bgc[c];
        float o;
        switch (p.mode_color) {
          case 0: o = f; break;
          case 1: o = __fadd_rn(b, f); break;
          case 2:
            if (!p.has_alpha) o = f;
            else if (p.alpha_premultiplied) o = __fadd_rn(f, __fmul_rn(b, one_m_fa));
            else o = __fmul_rn(__fadd_rn(__fmul_rn(f, fa), __fmul_rn(__fmul_rn(b, ba), one_m_fa)), rna);
            break;
          case 3: o = __fadd_rn(b, __fmul_rn(f, fa)); break;
          default: o = __fmul_rn(b, p.clamp ? fminf(fmaxf(f, 0.0f), 1.0f) : f); break;
        }
        oc[c] = o;
      }
      float oa = 1.0f;
      if (p.has_alpha) {
        const float f = __fmul_rn((float) fg[3], inv);
        switch (p.mode_alpha) {
          case 0: oa = f; break;
          case 1: oa = __fadd_rn(ba, f); break;
          case 2: oa = na; break;
          case 3: oa = ba; break;   // the alpha channel itself is carried over by kAlphaWeightedAdd
          default: oa = __fmul_rn(ba, p.clamp ? fa : f); break;
        }
      }
      of[0] = oc[0]; of[1] = oc[1]; of[2] = oc[2]; of[3] = p.has_alpha ? oa : 1.0f;
    }
  }
  (void) blended;
  if (p.save) p.save[(size_t) y * p.cw + x] = make_float4(of[0], of[1], of[2], of[3]);
  if (!p.out) return;
  // The picture handed out goes through libjxl's float pipeline when frames are blended, whose 8-bit conversion adds the
  // 32x32 blue-noise dither to EVERY channel, alpha included.  The stage runs over the frame's rectangle in 4-lane vectors
  // starting at its left edge and loads the pattern with one unaligned vector load at (y % 32) * 32 + (x % 32): lanes
  // that cross column 32 read on into the NEXT row of the table instead of wrapping.  Reproduced (bit-exact against the
  // reference on every blend mode, tests/test_gpu_composition.py).
  float dth = 0.0f;
  if (p.dither && kMax == 255) {
    const uint32_t o = p.orientation;
    if (o == 1) {
      const uint32_t left = p.x0 > 0 ? (uint32_t) p.x0 : 0u;
      const uint32_t start = x >= left ? left + ((x - left) & ~3u) : (x & ~3u);
      dth = p.dither[((y & 31) * 32 + (start & 31) + (x - start)) & 1023];
    } else {
      const bool flip_x = o == 2 || o == 3 || o == 7 || o == 8, flip_y = o == 3 || o == 4 || o == 6 || o == 7;
      const uint32_t ox = flip_x ? p.cw - 1 - x : x, oy = flip_y ? p.ch - 1 - y : y;
      dth = p.dither[(oy & 31) * 32 + (ox & 31)];
    }
  }
  uint32_t out[4];
  for (int c = 0; c < 4; ++c) out[c] = (uint32_t) lrintf(fminf(fmaxf(of[c] * (float) kMax + dth, 0.0f), (float) kMax));
  if (!p.has_alpha) out[3] = (uint32_t) kMax;
  uint8_t* dst = p.out + (size_t) y * p.canvas_stride + (size_t) x * sizeof(Pixel);
  if (sizeof(Pixel) == 4) *reinterpret_cast<uint32_t*>(dst) = out[0] | (out[1] << 8) | (out[2] << 16) | (out[3] << 24);
  else *reinterpret_cast<uint2*>(dst) = make_uint2(out[0] | (out[1] << 16), out[2] | (out[3] << 16));
}
}  // namespace

void LaunchComposite(const CompositeParams& p, cudaStream_t stream) {
  if (!p.cw || !p.ch) return;
  dim3 grid((p.cw + 255) / 256, p.ch, 1);
  if (p.bits16) CompositeKernel<uint2, 65535><<<grid, 256, 0, stream>>>(p);
  else CompositeKernel<uint32_t, 255><<<grid, 256, 0, stream>>>(p);
  ++g_launches_ac;
}

void LaunchPlace(const uint8_t* src, uint32_t src_stride, uint32_t fw, uint32_t fh, uint32_t bpp, int32_t x0, int32_t y0, uint32_t fill_alpha,
                 uint8_t* dst, uint32_t dst_stride, uint32_t cw, uint32_t ch, cudaStream_t stream) {
  if (!cw || !ch) return;
  dim3 grid((cw + 255) / 256, ch, 1);
  if (bpp == 8) PlaceKernel<uint2><<<grid, 256, 0, stream>>>(src, src_stride, fw, fh, x0, y0, make_uint2(0u, fill_alpha << 16), dst, dst_stride, cw, ch);
  else PlaceKernel<uint32_t><<<grid, 256, 0, stream>>>(src, src_stride, fw, fh, x0, y0, fill_alpha << 24, dst, dst_stride, cw, ch);
  ++g_launches_ac;
}

void LaunchUnsqueeze(const FrameDev& f, const SqStep* steps_host, cudaStream_t stream) {
  if (!f.sq_nch) return;
  if (f.sq_global) {
    SqueezeScatterGlobalKernel<<<f.sq_global, 256, 0, stream>>>(f);
    ++g_launches_ac;
  }
  for (uint32_t k = 0; k < f.sq_nsteps; ++k) {
    const uint32_t n = steps_host[k].horizontal ? steps_host[k].avg_h : steps_host[k].avg_w;
    SqueezeStepKernel<<<(n + 127) / 128, 128, 0, stream>>>(f, k);
    ++g_launches_ac;
  }
}

void LaunchOrient(const uint8_t* src, uint32_t src_stride, uint32_t w, uint32_t h, uint32_t bpp, uint32_t orientation, uint8_t* dst,
                  uint32_t dst_stride, cudaStream_t stream) {
  if (!w || !h) return;
  const uint32_t ow = orientation >= 5 ? h : w, oh = orientation >= 5 ? w : h;
  dim3 grid((ow + 255) / 256, oh, 1);
  if (bpp == 8) OrientKernel<uint2><<<grid, 256, 0, stream>>>(src, src_stride, w, h, orientation, dst, dst_stride);
  else OrientKernel<uint32_t><<<grid, 256, 0, stream>>>(src, src_stride, w, h, orientation, dst, dst_stride);
  ++g_launches_ac;
}

void LaunchColorMatrix(uint8_t* img, uint32_t stride, uint32_t width, uint32_t height, const ColorMatrixPlan* plan_dev, bool tonemap, bool bits16,
                       uint32_t* row_first, cudaStream_t stream) {
  if (!width || !height) return;
  const ColorMatrixTables16* t16 = reinterpret_cast<const ColorMatrixTables16*>(reinterpret_cast<const uint8_t*>(plan_dev) + ((sizeof(ColorMatrixPlan) + 255) & ~(size_t) 255));
  if (tonemap) {
    if (bits16) FirstBlackKernel<true><<<height, 256, 0, stream>>>(img, stride, width, t16->linearize, row_first);
    else FirstBlackKernel<false><<<height, 256, 0, stream>>>(img, stride, width, plan_dev->linearize, row_first);
    ++g_launches_ac;
  }
  dim3 grid((width + 255) / 256, std::min<uint32_t>(height, 1184), 1);
  if (bits16) ColorMatrix16Kernel<<<grid, 256, 0, stream>>>(img, stride, width, height, plan_dev, t16, row_first);
  else ColorMatrixKernel<<<grid, 256, 0, stream>>>(img, stride, width, height, plan_dev, row_first);
  ++g_launches_ac;
}

}  // namespace jxlb
