// Inverse VarDCT kernel (sm_100a): dequantisation + chroma-from-luma + LLF synthesis + all inverse transforms of the
// blocks contained in 64x64-pixel regions, quantised int16 coefficient planes -> XYB f32 planes.  This is the HBM
// roofline kernel of the path (18.25 B/pixel: 6 B coefficients + 0.25 B metadata read, 12 B written).
//
// Persistent CTAs (2 per SM, 256 threads each) walk the image's region list.  Per region:
//   * the 64 x 64 x 3 int16 coefficient tile (24 KB) arrives by ONE TMA tensor load (cp.async.bulk.tensor.3d,
//     out-of-bounds rows / columns zero-filled by the hardware) signalled on an mbarrier; the load of region n + 1 is
//     issued as soon as the dequantisation phase of region n has consumed the buffer, so it lands under the two IDCT
//     passes; the per-cell metadata and LF samples of region n + 1 travel in registers over the same interval;
//   * dequantisation reads its weights from shared memory (the matrices of every transform up to 32x32 -- 30 KB -- are
//     staged once per CTA) and the quant-bias adjustment from a 256-entry table per channel, branch-free; all-zero
//     8-coefficient units cost three 128-bit loads and six 128-bit stores;
//   * column pass and row pass keep a whole 1-D transform (8 .. 64 points) in the registers of one thread; each pass is
//     cut into 12 warp-sized items (channel, 32 columns, half of the rows) with the four luma items on warps of their
//     own; every warp votes on how many leading inputs are non-zero and runs a pruned transform when the high
//     frequencies are empty (always the case for the chroma channels of most pictures); the tile's row stride of 68
//     floats makes the column pass (scalar) and the row pass (128-bit) free of bank conflicts;
//   * finished rows leave with one 256-byte bulk async store each (cp.async.bulk shared -> global), so the 12 B/pixel
//     output is written in full lines by the copy engine while the CTA moves on.
// A warp-autonomous variant (one 32x32 area per warp, no CTA barriers) was measured and dropped: 9 warps per SM do not
// hide the dependent-issue latency of the transforms (184 us per 4096x4096 image against 135 us for this kernel).
// Blocks that are not contained in one region (larger than 64 pixels or straddling a region border) are left to
// ReconLargeKernel (kernels.cu), which runs after this kernel and overwrites their rectangles.
// Arithmetic: identical, operation by operation, to ReconRegion (recon.h), which tests/hostemu runs on the CPU.
#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "kernels.h"
#include "numeric_tables.h"
#include "recon.h"

namespace jxlb {

extern std::atomic<uint64_t> g_launches_ac;

namespace {

constexpr int kRT = 256;                       // threads per CTA: 2 dequantisation units each; 12 (channel, 32 columns / rows, half) items in the passes
constexpr int kTS = 68;                        // tile row stride (floats)
constexpr int kTP = kRegionDim * kTS;          // one channel of the tile
constexpr uint32_t kSmemTableFloats = 7680;    // quant tables 0 .. 10 (every transform up to 32x32): contiguous at pool offset 0
constexpr int kAdjHalf = 128;                  // quant-bias table covers q in [-128, 127]

struct ReconSmem {
  alignas(128) int16_t coef[3][kRegionDim][kRegionDim];  // TMA destination
  alignas(16) float tile[3 * kTP];
  float tables[kSmemTableFloats];
  float adj[3][2 * kAdjHalf];
  float lf[3][kRegionCells * kRegionCells];
  float llf_v[3][kRegionCells * kRegionCells];  // lowest-frequency coefficients from the LF image: [c][cell (oy + ky, ox + kx)] = coefficient (ky, kx)
  float llf[1 + 4 + 16 + 64];
  uint32_t cinfo[kRegionCells * kRegionCells];
  uint32_t dqoff[3][kRegionCells * kRegionCells];
  float cscale[kRegionCells * kRegionCells];
  float cfl[2];
  alignas(8) uint64_t mbar;
};

// cinfo bits
constexpr uint32_t kCiValid = 1u << 5;       // covered by a block that lies inside this region
constexpr uint32_t kCiSpecial = 1u << 6;     // one of the special 8x8 transforms
constexpr uint32_t kCiTransposed = 1u << 7;  // dequant matrix read transposed
// [0:4] strategy  [8:10] ox  [11:13] oy  [14:15] kcols_log - 3  [16:19] bx  [20:23] by

__device__ __forceinline__ uint32_t SmemAddr(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void MbarInit(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void MbarExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void MbarWait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(SmemAddr(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void TmaLoad3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(SmemAddr(dst)),
      "l"(map), "r"(SmemAddr(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}
__device__ __forceinline__ void BulkStoreRow(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(SmemAddr(ssrc)), "r"(bytes) : "memory");
}

// Idct1d<N> for an input whose entries K .. N - 1 are zero (they are not read): the same recursion with the additions of
// zeros and the products with zeros left out, so the result is the one Idct1d<N> gives (up to the sign of zeros).
template <int N, int K>
__device__ __forceinline__ void Idct1dK(float (&v)[N]) {
  if constexpr (K >= N) {
    Idct1d<N>(v);
  } else if constexpr (N == 2) {
    v[1] = v[0];
  } else {
    constexpr int KE = (K + 1) / 2, KO = K / 2;
    float e[N / 2], o[N / 2];
#pragma unroll
    for (int i = 0; i < KE; ++i) e[i] = v[2 * i];
#pragma unroll
    for (int i = 0; i < KO; ++i) o[i] = v[2 * i + 1];
    Idct1dK<N / 2, KE>(e);
    if constexpr (KO == 0) {
#pragma unroll
      for (int i = 0; i < N / 2; ++i) {
        v[i] = e[i];
        v[N - 1 - i] = e[i];
      }
    } else {
      constexpr int KO2 = KO + 1 < N / 2 ? KO + 1 : N / 2;
      if constexpr (KO < N / 2) o[KO] = o[KO - 1];
#pragma unroll
      for (int i = KO - 1; i > 0; --i) o[i] += o[i - 1];
      o[0] *= kSqrt2f;
      Idct1dK<N / 2, KO2>(o);
#pragma unroll
      for (int i = 0; i < N / 2; ++i) {
        const float m = WcMul<N>(i) * o[i];
        v[i] = e[i] + m;
        v[N - 1 - i] = e[i] - m;
      }
    }
  }
}

// OR of the bit patterns of v[A .. B): zero iff all of them are +-0.
template <int A, int B, int N>
__device__ __forceinline__ uint32_t OrBits(const float (&v)[N]) {
  uint32_t r = 0;
#pragma unroll
  for (int i = A; i < B; ++i) r |= __float_as_uint(v[i]);
  return r & 0x7FFFFFFFu;
}

// Picks the cheapest transform the warp's inputs allow.  High frequencies are zero for most blocks (and for the whole X
// and B channels of most pictures); the lanes of a warp that reach this point together work on neighbouring columns (or
// rows) of the same blocks, so the choice is made per warp (no divergence) from a vote over the lanes.
template <int N>
__device__ __forceinline__ void IdctAdaptive(float (&v)[N]) {
  if constexpr (N < 16) {
    Idct1d<N>(v);
  } else if constexpr (N > 32) {
    const unsigned m = __activemask();
    if (__any_sync(m, OrBits<N / 2, N>(v) != 0)) Idct1d<N>(v);
    else Idct1dK<N, N / 2>(v);
  } else {
    const unsigned m = __activemask();
    const bool hi = OrBits<N / 2, N>(v) != 0, mid = OrBits<N / 4, N / 2>(v) != 0, lo = OrBits<N / 8, N / 4>(v) != 0;
    const unsigned vh = __ballot_sync(m, hi), vm = __ballot_sync(m, mid), vl = __ballot_sync(m, lo);
    if (vh) Idct1d<N>(v);
    else if (vm) Idct1dK<N, N / 2>(v);
    else if (vl) Idct1dK<N, N / 4>(v);
    else Idct1dK<N, N / 8>(v);
  }
}

// Column pass of one block column.  p: top of the column in the tile.  llf_col: when non-null, this column carries the
// block's lowest frequencies in its first N / 8 rows (one per cell row, llf_col[ky * 8]), which replace what the
// dequantisation left there.
template <int N>
__device__ __forceinline__ void ColumnIdct(float* p, const float* llf_col) {
  constexpr int BY = N / 8;
  float v[N];
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = p[i * kTS];
  if (llf_col) {
#pragma unroll
    for (int ky = 0; ky < BY; ++ky) v[ky] = llf_col[ky * kRegionCells];
  }
  IdctAdaptive<N>(v);
#pragma unroll
  for (int i = 0; i < N; ++i) p[i * kTS] = v[i];
}
template <int N>
__device__ __forceinline__ void RowIdct(float* p) {  // p: 16-byte aligned start of the block's row in the tile
  float v[N];
#pragma unroll
  for (int i = 0; i < N; i += 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + i);
    v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
  }
  IdctAdaptive<N>(v);
#pragma unroll
  for (int i = 0; i < N; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}

struct RegionPrefetch {  // metadata of the next region, held in registers while the current one is processed
  uint32_t strategy, off, hfmul;  // threads 0 .. 63: one cell each
  float lf;                       // every thread: one LF sample (channel tid / 64, cell tid % 64)
  float kx, kb;                   // thread 64
};

__device__ __forceinline__ void LoadRegionMeta(const FrameDev& f, uint32_t rx, uint32_t ry, int tid, RegionPrefetch* m) {
  const uint32_t cx0 = rx * kRegionCells, cy0 = ry * kRegionCells;
  const int ci = tid & 63;
  const uint32_t gx = cx0 + (uint32_t) (ci & 7), gy = cy0 + (uint32_t) (ci >> 3);
  const bool inside = gx < f.w8 && gy < f.h8;
  m->strategy = 0xFF;
  m->off = 0;
  m->hfmul = 1;
  if (tid < 64 && inside) {
    const size_t o = (size_t) gy * f.w8 + gx;
    m->strategy = f.cell_strategy[o];
    m->off = f.cell_off[o];
    m->hfmul = f.cell_hfmul[o];
  }
  const int c = tid >> 6;
  m->lf = (inside && c < 3) ? f.lf[(size_t) c * f.h8 * f.lf_stride + (size_t) gy * f.lf_stride + gx] : 0.0f;
  if (tid == 64) {
    // CfL factors of the region's 64x64 tile (a contained block's top-left corner lies in it)
    const size_t t = (size_t) ry * f.w64 + rx;
    m->kx = f.cfl.base_x + (float) f.xfromy[t] / (float) f.cfl.colour_factor;
    m->kb = f.cfl.base_b + (float) f.bfromy[t] / (float) f.cfl.colour_factor;
  }
}

__device__ __forceinline__ void StoreRegionMeta(const FrameDev& f, const NumericTables& nt, ReconSmem& sh, int tid, const RegionPrefetch& m, uint32_t rx,
                                                uint32_t ry, bool list_large) {
  if (tid < 64) {
    const int ix = tid & 7, iy = tid >> 3;
    uint32_t info = 0, d0 = 0, d1 = 0, d2 = 0;
    float scale = 0.0f;
    if (m.strategy != 0xFF) {
      const uint32_t t = m.strategy & 0x7Fu;
      const int dx = (int) (m.off & 0xFF), dy = (int) (m.off >> 8);
      const int ox = ix - dx, oy = iy - dy;
      const int bx = (int) StrategyCellsX(t), by = (int) StrategyCellsY(t);
      const bool contained = ox >= 0 && oy >= 0 && ox + bx <= kRegionCells && oy + by <= kRegionCells;
      if (!contained && (m.strategy & 0x80u) && list_large) {  // top-left cell of a block no region contains: left to ReconLargeListKernel
        const uint32_t slot = atomicAdd(f.large_list, 1u);
        f.large_list[1 + slot] = (rx * kRegionCells + (uint32_t) ix) | ((ry * kRegionCells + (uint32_t) iy) << 16);
      }
      if (contained) {
        const uint32_t qt = StrategyQuantTable(t);
        const uint32_t kl = 3u + (uint32_t) FloorLog2((uint32_t) (bx > by ? bx : by));
        info = t | kCiValid | (IsSpecial8x8(t) ? kCiSpecial : 0u) | ((by >= bx && !nt.dequant_symmetric[qt]) ? kCiTransposed : 0u) |
               ((uint32_t) ox << 8) | ((uint32_t) oy << 11) | ((kl - 3u) << 14) | ((uint32_t) bx << 16) | ((uint32_t) by << 20);
        d0 = nt.dequant_off[qt][0];
        d1 = nt.dequant_off[qt][1];
        d2 = nt.dequant_off[qt][2];
        scale = 65536.0f / (float) f.global_scale / (float) m.hfmul;
      }
    }
    sh.cinfo[tid] = info;
    sh.dqoff[0][tid] = d0;
    sh.dqoff[1][tid] = d1;
    sh.dqoff[2][tid] = d2;
    sh.cscale[tid] = scale;
  }
  if (tid < 192) sh.lf[tid >> 6][tid & 63] = m.lf;
  if (tid == 64) {
    sh.cfl[0] = m.kx;
    sh.cfl[1] = m.kb;
  }
}

// Dequantises the 8 coefficients of one (unit, channel): out[j] = adj(q_j) * (w_j * s), libjxl's order.  Branch-free for
// |q| < 128 (adj[] maps 0 to 0, so zero coefficients need no test); units holding a larger value take the exact formula.
__device__ __forceinline__ void DequantUnit(const uint4 raw, const float* adj, const uint32_t* rcp11, uint32_t c, const float* tab, bool tab_shared,
                                            uint32_t off, uint32_t kstep, float s, float (&out)[8]) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
  // any halfword outside [-128, 127]?  (h + 128) has bits above bit 7 set
  uint32_t big = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) big |= (((w[k] & 0xFFFFu) + 0x80u) & 0xFFFFu) | (((w[k] >> 16) + 0x80u) & 0xFFFFu);
  float wt[8];
  if (kstep == 1) {  // consecutive weights, 32-byte aligned (off is a multiple of 8)
    float4 a, b;
    if (tab_shared) {
      a = *reinterpret_cast<const float4*>(tab + off);
      b = *reinterpret_cast<const float4*>(tab + off + 4);
    } else {
      a = __ldg(reinterpret_cast<const float4*>(tab + off));
      b = __ldg(reinterpret_cast<const float4*>(tab + off + 4));
    }
    wt[0] = a.x; wt[1] = a.y; wt[2] = a.z; wt[3] = a.w; wt[4] = b.x; wt[5] = b.y; wt[6] = b.z; wt[7] = b.w;
  } else if (tab_shared) {
#pragma unroll
    for (int j = 0; j < 8; ++j) wt[j] = tab[off + (uint32_t) j * kstep];
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) wt[j] = __ldg(tab + off + (uint32_t) j * kstep);
  }
  if (!(big & 0xFF00u)) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t idx = ((w[j >> 1] ^ 0x00800080u) >> ((j & 1) * 16)) & 0xFFu;   // (q + 128) for q in [-128, 127]
      out[j] = __fmul_rn(adj[idx], __fmul_rn(wt[j], s));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int q = (int) (int16_t) (w[j >> 1] >> ((j & 1) * 16));
      out[j] = q ? __fmul_rn(AdjustQuantBiasRcp(q, c, rcp11), __fmul_rn(wt[j], s)) : 0.0f;
    }
  }
}

__global__ void __launch_bounds__(kRT, 2) ReconRegionTmaKernel(const FrameDev f, const NumericTables* ntp, const __grid_constant__ CUtensorMap coef_map,
                                                               uint32_t nrx, uint32_t nregions) {
  if (*f.frame_bad) return;
  extern __shared__ __align__(128) uint8_t recon_smem_raw[];
  ReconSmem& sh = *reinterpret_cast<ReconSmem*>(recon_smem_raw);
  const NumericTables& nt = *ntp;
  const int tid = (int) threadIdx.x;
  uint32_t r = blockIdx.x;
  if (r >= nregions) return;
  // ---- once per CTA: barrier, tables
  if (tid == 0) {
    MbarInit(&sh.mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (uint32_t i = (uint32_t) tid; i < kSmemTableFloats; i += kRT) sh.tables[i] = __ldg(nt.dequant + i);
  for (int i = tid; i < 3 * 2 * kAdjHalf; i += kRT) {
    const int c = i / (2 * kAdjHalf), q = i % (2 * kAdjHalf) - kAdjHalf;
    sh.adj[c][q + kAdjHalf] = AdjustQuantBiasRcp(q, (uint32_t) c, nt.rcp11);
  }
  for (int i = tid; i < 85; i += kRT) {
    const int l = i < 1 ? 0 : i < 5 ? 1 : i < 21 ? 2 : 3;
    sh.llf[i] = nt.llf[l][i - LlfSharedOffset(l)];
  }
  const float xqm = QmScale(f.x_qm_scale), bqm = QmScale(f.b_qm_scale);
  const size_t pplane = (size_t) f.plane_h * f.plane_stride;
  RegionPrefetch meta;
  LoadRegionMeta(f, r % nrx, r / nrx, tid, &meta);
  __syncthreads();  // barrier initialised, tables staged
  if (tid == 0) {
    MbarExpectTx(&sh.mbar, (uint32_t) sizeof(sh.coef));
    TmaLoad3d(&sh.coef[0][0][0], &coef_map, &sh.mbar, (int) ((r % nrx) * kRegionDim), (int) ((r / nrx) * kRegionDim), 0);
  }
  uint32_t parity = 0;
  while (r < nregions) {
    const uint32_t rx = r % nrx, ry = r / nrx;
    const uint32_t rn = r + gridDim.x;
    // ---- P0: this region's metadata from registers to shared memory; start fetching the next region's
    // (the previous region's rows must have left the tile before P1 overwrites it)
    StoreRegionMeta(f, nt, sh, tid, meta, rx, ry, true);
    if (rn < nregions) LoadRegionMeta(f, rn % nrx, rn / nrx, tid, &meta);
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();
    // lowest frequencies of every block from the LF image (one coefficient per thread, computed while the coefficient
    // tile is still in flight; the column pass puts them in place):
    //   coefficient (ky, kx) = sum_ny A_by[ky][ny] * (sum_nx A_bx[kx][nx] * LF[oy + ny][ox + nx])
    if (tid < 192) {
      const int c = tid >> 6, ci = tid & 63;
      const uint32_t info = sh.cinfo[ci];
      float v = 0.0f;
      if (info & kCiValid) {
        const int ox = (int) ((info >> 8) & 7u), oy = (int) ((info >> 11) & 7u), bx = (int) ((info >> 16) & 15u), by = (int) ((info >> 20) & 15u);
        const float* lf = sh.lf[c] + oy * kRegionCells + ox;
        if (bx == 1 && by == 1) {
          v = lf[0];
        } else {
          const float* ay = sh.llf + LlfSharedOffset(FloorLog2((uint32_t) by)) + ((ci >> 3) - oy) * by;
          const float* ax = sh.llf + LlfSharedOffset(FloorLog2((uint32_t) bx)) + ((ci & 7) - ox) * bx;
          for (int ny = 0; ny < by; ++ny) {
            const float* l = lf + ny * kRegionCells;
            float rowacc = 0.0f;
            if (bx == 4) {
              rowacc += ax[0] * l[0]; rowacc += ax[1] * l[1]; rowacc += ax[2] * l[2]; rowacc += ax[3] * l[3];
            } else if (bx == 2) {
              rowacc += ax[0] * l[0]; rowacc += ax[1] * l[1];
            } else {
              for (int nx = 0; nx < bx; ++nx) rowacc += ax[nx] * l[nx];
            }
            v += ay[ny] * rowacc;
          }
        }
      }
      sh.llf_v[c][ci] = v;
    }
    MbarWait(&sh.mbar, parity);
    parity ^= 1;
    // ---- P1: dequantise + chroma-from-luma (+ the lowest frequencies from the LF image) into the tile
    const float kx = sh.cfl[0], kb = sh.cfl[1];
    for (int u = tid; u < kRegionDim * kRegionDim / 8; u += kRT) {
      const int row = u >> 3, cg = u & 7;
      const int ci = (row >> 3) * kRegionCells + cg;
      const uint32_t info = sh.cinfo[ci];
      if (!(info & kCiValid)) continue;
      const uint4 rx4 = *reinterpret_cast<const uint4*>(&sh.coef[0][row][cg * 8]);
      const uint4 ry4 = *reinterpret_cast<const uint4*>(&sh.coef[1][row][cg * 8]);
      const uint4 rb4 = *reinterpret_cast<const uint4*>(&sh.coef[2][row][cg * 8]);
      const int ox = (int) ((info >> 8) & 7u), oy = (int) ((info >> 11) & 7u);
      const uint32_t pr = (uint32_t) (row - oy * 8), pc0 = (uint32_t) ((cg - ox) * 8);
      const int bx = (int) ((info >> 16) & 15u), by = (int) ((info >> 20) & 15u);
      const bool zx = (rx4.x | rx4.y | rx4.z | rx4.w) == 0, zy = (ry4.x | ry4.y | ry4.z | ry4.w) == 0, zb = (rb4.x | rb4.y | rb4.z | rb4.w) == 0;
      float vx[8], vy[8], vb[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) vx[j] = vy[j] = vb[j] = 0.0f;
      if (!(zx && zy && zb)) {
        const uint32_t kl = 3u + ((info >> 14) & 3u);
        const bool tr = (info & kCiTransposed) != 0;
        const uint32_t k0 = tr ? ((pc0 << kl) + pr) : ((pr << kl) + pc0);
        const uint32_t kstep = tr ? (1u << kl) : 1u;
        const float sy = sh.cscale[ci], sx = sy * xqm, sb = sy * bqm;
        const uint32_t o0 = sh.dqoff[0][ci] + k0, o1 = sh.dqoff[1][ci] + k0, o2 = sh.dqoff[2][ci] + k0;
        const bool in_smem = o2 + 7u * kstep < kSmemTableFloats;  // tables are ordered X, Y, B inside one quant table
        const float* tab = in_smem ? sh.tables : nt.dequant;
        if (!zy) DequantUnit(ry4, sh.adj[1], nt.rcp11, 1, tab, in_smem, o1, kstep, sy, vy);
        if (!zx) DequantUnit(rx4, sh.adj[0], nt.rcp11, 0, tab, in_smem, o0, kstep, sx, vx);
        if (!zb) DequantUnit(rb4, sh.adj[2], nt.rcp11, 2, tab, in_smem, o2, kstep, sb, vb);
        if (!zy) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            vx[j] = __fadd_rn(__fmul_rn(kx, vy[j]), vx[j]);
            vb[j] = __fadd_rn(__fmul_rn(kb, vy[j]), vb[j]);
          }
        }
      }
      float* t = sh.tile + row * kTS + cg * 8;
      *reinterpret_cast<float4*>(t) = make_float4(vx[0], vx[1], vx[2], vx[3]);
      *reinterpret_cast<float4*>(t + 4) = make_float4(vx[4], vx[5], vx[6], vx[7]);
      *reinterpret_cast<float4*>(t + kTP) = make_float4(vy[0], vy[1], vy[2], vy[3]);
      *reinterpret_cast<float4*>(t + kTP + 4) = make_float4(vy[4], vy[5], vy[6], vy[7]);
      *reinterpret_cast<float4*>(t + 2 * kTP) = make_float4(vb[0], vb[1], vb[2], vb[3]);
      *reinterpret_cast<float4*>(t + 2 * kTP + 4) = make_float4(vb[4], vb[5], vb[6], vb[7]);
    }
    __syncthreads();
    // the coefficient buffer is free: fetch the next region's tile under the two transform passes
    if (tid == 0 && rn < nregions) {
      MbarExpectTx(&sh.mbar, (uint32_t) sizeof(sh.coef));
      TmaLoad3d(&sh.coef[0][0][0], &coef_map, &sh.mbar, (int) ((rn % nrx) * kRegionDim), (int) ((rn / nrx) * kRegionDim), 0);
    }
    // ---- P3: special 8x8 transforms (disjoint from the DCT blocks) and the column pass.
    // The passes are split into 12 items (channel, 32 columns, upper / lower half of the rows) of one warp each; the four Y
    // items go to warps of their own: in most pictures the chroma channels have nothing but their lowest frequencies, so
    // their (pruned) transforms cost a third of Y's, and a static "channel per warp pair" split would leave four of six
    // warps waiting at the barrier.
    const int warp = tid >> 5, lane = tid & 31;
    // items of this warp: c | g << 2 | h << 3, 0xFF = none
    const uint32_t items = warp < 4 ? (0xFF00u | (1u | ((uint32_t) (warp >> 1) << 2) | ((uint32_t) (warp & 1) << 3)))
                                    : ((((uint32_t) (warp & 2)) | ((uint32_t) (warp & 1) << 2) | 8u) << 8) | (((uint32_t) (warp & 2)) | ((uint32_t) (warp & 1) << 2));
    if (tid < 192) {
      const int c = tid >> 6, ci = tid & 63;
      const uint32_t info = sh.cinfo[ci];
      if ((info & (kCiValid | kCiSpecial)) == (kCiValid | kCiSpecial)) {
        float* rect = sh.tile + c * kTP + (ci >> 3) * 8 * kTS + (ci & 7) * 8;
        rect[0] = sh.llf_v[c][ci];
        SpecialTransform8x8(info & 31u, rect, kTS, nt.afv_basis);
      }
    }
    __syncwarp();
#pragma unroll 1
    for (int k = 0; k < 2; ++k) {
      const uint32_t it = (items >> (8 * k)) & 0xFFu;
      if (it == 0xFFu) break;
      const int c = (int) (it & 3u), x = (int) ((it >> 2) & 1u) * 32 + lane, h = (int) (it >> 3);
      float* col = sh.tile + c * kTP + x;
      for (int iy = 4 * h; iy < 4 * h + 4;) {
        const uint32_t inf = sh.cinfo[iy * kRegionCells + (x >> 3)];
        const int by = (int) ((inf >> 20) & 15u);
        if ((inf & (kCiValid | kCiSpecial)) != kCiValid || (int) ((inf >> 11) & 7u) != iy) {
          ++iy;
          continue;
        }
        float* p = col + iy * 8 * kTS;
        const int ox = (int) ((inf >> 8) & 7u), bx = (int) ((inf >> 16) & 15u);
        const int kxi = x - ox * 8;
        const float* llf_col = kxi < bx ? sh.llf_v[c] + iy * kRegionCells + ox + kxi : nullptr;
        switch (by) {
          case 1: ColumnIdct<8>(p, llf_col); break;
          case 2: ColumnIdct<16>(p, llf_col); break;
          case 4: ColumnIdct<32>(p, llf_col); break;
          default: ColumnIdct<64>(p, llf_col); break;
        }
        iy += by;
      }
    }
    __syncthreads();
    // ---- P4: row pass, same items with rows and columns exchanged
#pragma unroll 1
    for (int k = 0; k < 2; ++k) {
      const uint32_t it = (items >> (8 * k)) & 0xFFu;
      if (it == 0xFFu) break;
      const int c = (int) (it & 3u), y = (int) ((it >> 2) & 1u) * 32 + lane, h = (int) (it >> 3);
      float* rowp = sh.tile + c * kTP + y * kTS;
      for (int ix = 4 * h; ix < 4 * h + 4;) {
        const uint32_t inf = sh.cinfo[(y >> 3) * kRegionCells + ix];
        const int bx = (int) ((inf >> 16) & 15u);
        if ((inf & (kCiValid | kCiSpecial)) != kCiValid || (int) ((inf >> 8) & 7u) != ix) {
          ++ix;
          continue;
        }
        float* p = rowp + ix * 8;
        switch (bx) {
          case 1: RowIdct<8>(p); break;
          case 2: RowIdct<16>(p); break;
          case 4: RowIdct<32>(p); break;
          default: RowIdct<64>(p); break;
        }
        ix += bx;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this thread's tile writes -> visible to the copy engine
    __syncthreads();
    // every row leaves with one 256-byte bulk store
    if (tid < 192) {
      const int c = tid >> 6, y = tid & 63;
      const uint32_t gy = ry * kRegionDim + (uint32_t) y;
      if (gy < f.plane_h) {
        BulkStoreRow(f.xyb0 + (size_t) c * pplane + (size_t) gy * f.plane_stride + rx * kRegionDim, sh.tile + c * kTP + y * kTS, kRegionDim * 4);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    r = rn;
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn GetEncodeTiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

}  // namespace

// Returns false when the TMA path cannot be used for this frame (no driver entry point, unaligned planes): the caller
// then runs the plain-load kernel of kernels.cu.
bool LaunchReconTma(const FrameDev& f, const NumericTables* nt_dev, cudaStream_t stream) {
  static const bool disabled = getenv("JXLB_NO_TMA_RECON") != nullptr;
  if (disabled) return false;
  EncodeTiledFn enc = GetEncodeTiled();
  if (!enc) return false;
  // the kernels stage pool[0, kSmemTableFloats) = quant tables 0 .. 10 (X, Y, B each) in shared memory
  static const bool layout_ok = GetHostNumericTables().tables.dequant_off[0][0] == 0 && GetHostNumericTables().tables.dequant_off[11][0] == kSmemTableFloats;
  if (!layout_ok) return false;
  if ((reinterpret_cast<uintptr_t>(f.coef) & 15) || (f.coef_stride & 7) || (f.plane_stride & 63) || (f.plane_h & 63) ||
      (reinterpret_cast<uintptr_t>(f.xyb0) & 15))
    return false;
  CUtensorMap map64;
  const cuuint64_t dims[3] = {f.coef_stride, f.coef_h, 3};
  const cuuint64_t strides[2] = {(cuuint64_t) f.coef_stride * 2, (cuuint64_t) f.coef_stride * 2 * f.coef_h};
  const cuuint32_t box64[3] = {kRegionDim, kRegionDim, 3};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (enc(&map64, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, f.coef, dims, strides, box64, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  static int region_ctas = 0;
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(ReconRegionTmaKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(ReconSmem));
    cudaFuncSetAttribute(ReconRegionTmaKernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ReconRegionTmaKernel, kRT, sizeof(ReconSmem));
    if (per_sm < 1) per_sm = 1;
    region_ctas = sms * per_sm;
  });
  const uint32_t nrx = (f.w8 + kRegionCells - 1) / kRegionCells, nry = (f.h8 + kRegionCells - 1) / kRegionCells;
  const uint32_t nregions = nrx * nry;
  const uint32_t grid = nregions < (uint32_t) region_ctas ? nregions : (uint32_t) region_ctas;
  ReconRegionTmaKernel<<<grid, kRT, sizeof(ReconSmem), stream>>>(f, nt_dev, map64, nrx, nregions);
  ++g_launches_ac;
  return true;
}

}  // namespace jxlb
