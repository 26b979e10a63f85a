// CUDA kernels of the JPEG XL decode path (sm_100a).
//
// Stream kernels (entropy-coded sections): every section of a frame is an inherently serial bitstream (ANS state,
// context depends on previously decoded symbols), so the unit of parallelism is the section: (image, LF group) and
// (image, group).  This first version gives each section its own warp with one active lane running the shared
// host/device section decoder; a batch of 64 4096x4096 images has 16 384 group sections + 256 LF-group sections in
// flight.  Numeric kernels: LF dequant + smoothing per cell, dequant/CfL/LLF/inverse-VarDCT per 64x64 region in
// shared memory (recon.h), per-pixel filters and colour conversion (pixel_stages.h).
#include "kernels.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "recon.h"

namespace jxlb {

namespace {

std::atomic<uint64_t> g_launches{0};

constexpr int kStreamBlockThreads = 32;  // one warp per section, lane 0 decodes
// Shared memory of an LF-group stream: only the modular decoder's rows and LUTs (17 KB).  The stream's entropy code
// (context map + alias tables, 15-50 KB) is read through L1 instead (an L1 hit costs the same ~30 cycles as LDS), so
// that LF-group CTAs of one batch stay co-resident with the 180 KB AC CTAs and the reconstruction CTAs of another.
constexpr uint32_t kLfFastCodeBytes = 0;
constexpr uint32_t kLfFastInts = 3968;  // >= ModFastScratch::Ints(256) = 3948 (an LF group is at most 256 cells wide)

struct SyncThreads {
  __device__ void operator()() const { __syncthreads(); }
};

__device__ StreamScratch CarveScratch(const ScratchLayout& s, uint32_t job, uint8_t** hf_arena, uint32_t** perm) {
  uint8_t* p = s.base + (uint64_t) job * s.bytes_per_job;
  StreamScratch sc;
  sc.arena.Init(p, s.arena_bytes);
  p += (s.arena_bytes + 255u) & ~255u;
  sc.wp = reinterpret_cast<int32_t*>(p);
  sc.wp_ints = s.wp_ints;
  p += ((uint64_t) s.wp_ints * 4 + 255u) & ~(uint64_t) 255u;
  sc.nzmap = p;
  p += 3 * 1024;
  sc.lz77 = nullptr;
  sc.lz77_mask = 0;
  if (hf_arena) *hf_arena = p;
  p += (s.hf_arena_bytes + 255u) & ~255u;
  if (perm) *perm = reinterpret_cast<uint32_t*>(p);
  return sc;
}

__device__ void BindLz77Window(const ScratchLayout& s, const StreamJob& job, StreamScratch* sc) {
  if (job.lz_slot && s.lz_base) {
    sc->lz77 = reinterpret_cast<uint32_t*>(s.lz_base + (uint64_t) (job.lz_slot - 1) * s.lz_entries * 4u);
    sc->lz77_mask = s.lz_entries - 1;
  }
}

__global__ void __launch_bounds__(kStreamBlockThreads) SingleSectionKernel(const FrameDev* frames, const StreamJob* jobs,
                                                                           uint32_t njobs, NaturalOrders nat, ScratchLayout scratch) {
  const uint32_t j = blockIdx.x;
  if (j >= njobs || threadIdx.x != 0) return;
  const StreamJob job = jobs[j];
  const FrameDev& f = frames[job.frame];
  uint8_t* hf_mem;
  uint32_t* perm;
  StreamScratch sc = CarveScratch(scratch, j, &hf_mem, &perm);
  BindLz77Window(scratch, job, &sc);
  Arena hf;
  hf.Init(hf_mem, scratch.hf_arena_bytes);
  f.status[job.status_slot] = DecodeSingleSectionFrame(f, nat, sc, hf, perm, scratch.max_local_nodes);
}

// Block placement of one LF group by a whole warp (the serial loop at the end of DecodeLfGroupSection, same results and
// the same error conditions).  "The next BlockInfo entry goes to the first uncovered cell in raster order" is a serial
// rule, but only over the ~4 K blocks, not the 64 K cells: lane 0 walks a coverage bitmap in shared memory (one bit per
// cell, find-first-set per 32 cells) and records each block's position; the per-cell planes are then filled by all lanes.
__device__ int PlaceBlocksWarp(const FrameDev& f, uint32_t lfg, uint32_t* bitmap /* 256 rows x 8 words, shared */) {
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t cx0, cy0, w8, h8, w64, h64;
  LfGroupRect(f, lfg, &cx0, &cy0, &w8, &h8, &w64, &h64);
  const uint32_t nb = f.nb_blocks[lfg];
  int32_t* binfo = f.blockinfo + f.blockinfo_off[lfg];
  // sharpness plane -> bytes (range-checked); coverage bitmap with the cells right of the group marked as covered
  bool bad = false;
  for (uint32_t i = lane; i < w8 * h8; i += 32) {
    const uint32_t y = i / w8, x = i - y * w8;
    const int32_t sh = f.sharpness_i32[(size_t) (cy0 + y) * f.lf_stride + cx0 + x];
    if (sh < 0 || sh > 7) bad = true;
    f.cell_sharp[(size_t) (cy0 + y) * f.w8 + cx0 + x] = (uint8_t) sh;
  }
  for (uint32_t i = lane; i < h8 * 8; i += 32) {
    const uint32_t base = (i & 7u) * 32u;
    bitmap[i] = base >= w8 ? 0xFFFFFFFFu : (w8 - base >= 32u ? 0u : (0xFFFFFFFFu << (w8 - base)));
  }
  __syncwarp();
  int st = kOk;
  if (lane == 0) {
    const uint32_t nwords = h8 * 8;
    uint32_t pos = 0;
    for (uint32_t k = 0; k < nb; ++k) {
      uint32_t free_bits = 0;
      while (pos < nwords && (free_bits = ~bitmap[pos]) == 0) ++pos;
      if (pos >= nwords) {  // blocks left over after every cell is covered
        st = kErrBadStream;
        break;
      }
      const uint32_t x = (pos & 7u) * 32u + (uint32_t) __ffs((int) free_bits) - 1u, y = pos >> 3;
      const int32_t t = binfo[k];
      int32_t q = binfo[nb + k];
      q = 1 + (q < 0 ? 0 : q > 255 ? 255 : q);  // hf_mul is clamped to [1, 256]
      if (t < 0 || t >= kNumStrategies) {
        st = kErrBadStream;
        break;
      }
      const uint32_t bx = StrategyCellsX((uint32_t) t), by = StrategyCellsY((uint32_t) t);
      // inside the LF group, and not straddling a 256x256 group boundary (so never straddling a bitmap word either)
      if (x + bx > w8 || y + by > h8 || (x % kGroupCells) + bx > kGroupCells || (y % kGroupCells) + by > kGroupCells) {
        st = kErrBadStream;
        break;
      }
      const uint32_t mask = (bx >= 32u ? 0xFFFFFFFFu : ((1u << bx) - 1u)) << (x & 31u);
      bool overlap = false;
      for (uint32_t yy = 0; yy < by; ++yy) {
        uint32_t& wd = bitmap[(y + yy) * 8 + (x >> 5)];
        if (wd & mask) overlap = true;
        wd |= mask;
      }
      if (overlap) {
        st = kErrBadStream;
        break;
      }
      binfo[k] = t | (int32_t) (x << 8) | (int32_t) (y << 16);
      binfo[nb + k] = q;
    }
    if (st == kOk) {  // a cell left uncovered means BlockInfo ran out of blocks
      for (; pos < nwords; ++pos)
        if (~bitmap[pos] != 0) {
          st = kErrBadStream;
          break;
        }
    }
  }
  st = __shfl_sync(0xFFFFFFFFu, st, 0);
  if (__any_sync(0xFFFFFFFFu, bad)) st = st == kOk ? kErrBadStream : st;
  if (st != kOk) return st;
  for (uint32_t k = lane; k < nb; k += 32) {
    const uint32_t rec = (uint32_t) binfo[k];
    const uint32_t t = rec & 0xFFu, x = (rec >> 8) & 0xFFu, y = rec >> 16;
    const uint16_t q = (uint16_t) binfo[nb + k];
    const uint32_t bx = StrategyCellsX(t), by = StrategyCellsY(t);
    for (uint32_t yy = 0; yy < by; ++yy) {
      const size_t o = (size_t) (cy0 + y + yy) * f.w8 + cx0 + x;
      for (uint32_t xx = 0; xx < bx; ++xx) {
        f.cell_strategy[o + xx] = (uint8_t) ((yy | xx) == 0 ? (t | 0x80u) : t);
        f.cell_hfmul[o + xx] = q;
        f.cell_off[o + xx] = (uint16_t) ((yy << 8) | xx);
      }
    }
  }
  return kOk;
}

// Several LF groups per CTA, one warp each.  The CTA is sized to take a whole SM for itself (200 KB of shared memory): the LF stage is a set of serial dependency
// chains that leave an SM's issue slots almost idle, and when its warps were spread one or two per SM over the whole GPU
// they competed for issue slots with the dense kernels of the other batches in flight -- the chains ran 45 % slower and
// the dense kernels lost throughput too.  Packed, a 64-image batch's 256 chains hold 22 SMs at ~3 warps per scheduler
// (each warp issues once every ~3 cycles: the schedulers are full) and the other 126 SMs run the dense kernels undisturbed.
constexpr int kLfMaxWarps = 12;
__global__ void __launch_bounds__(kLfMaxWarps * 32, 1) LfGroupKernel(const FrameDev* frames, const StreamJob* jobs, uint32_t njobs,
                                                                     ScratchLayout scratch, int warp_place) {
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t j = blockIdx.x * (blockDim.x >> 5) + warp;
  if (j >= njobs) return;
  const StreamJob job = jobs[j];
  const FrameDev& f = frames[job.frame];
  extern __shared__ __align__(16) uint8_t lf_smem_all[];
  uint8_t* lf_smem = lf_smem_all + (size_t) warp * (kLfFastCodeBytes + kLfFastInts * 4);
  int st = kOk;
  if (lane == 0) {
    StreamScratch sc = CarveScratch(scratch, j, nullptr, nullptr);
    sc.fast = lf_smem;
    sc.fast_code_bytes = kLfFastCodeBytes;
    sc.fast_ints = kLfFastInts;
    BitReader br;
    const uint32_t sec = 1 + job.index;
    br.Init(f.cs, f.cs_bytes, f.sec_bit_begin[sec], f.sec_bit_end[sec]);
    st = DecodeLfGroupSection(br, f, job.index, sc, scratch.max_local_nodes, /*place_blocks=*/!warp_place);
  }
  __syncwarp();
  st = __shfl_sync(0xFFFFFFFFu, st, 0);
  if (st == kOk && warp_place) st = PlaceBlocksWarp(f, job.index, reinterpret_cast<uint32_t*>(lf_smem));
  if (lane == 0) f.status[job.status_slot] = st;
}

__global__ void __launch_bounds__(kStreamBlockThreads) PassGroupKernel(const FrameDev* frames, const StreamJob* jobs, uint32_t njobs,
                                                                       NaturalOrders nat, ScratchLayout scratch) {
  const uint32_t j = blockIdx.x;
  if (j >= njobs || threadIdx.x != 0) return;
  const StreamJob job = jobs[j];
  const FrameDev& f = frames[job.frame];
  StreamScratch sc = CarveScratch(scratch, j, nullptr, nullptr);
  BindLz77Window(scratch, job, &sc);
  int st = kOk;
  // progressive frames: the group's sections of pass 0, 1, ... in turn (their coefficients add up)
  for (uint32_t pass = 0; pass < f.num_passes && st == kOk; ++pass) {
    BitReader br;
    const uint32_t sec = 1 + f.num_lf_groups + 1 + pass * f.num_groups + job.index;
    br.Init(f.cs, f.cs_bytes, f.sec_bit_begin[sec], f.sec_bit_end[sec]);
    if (f.encoding == 0) {
      // the group's blocks come from its LF group's placement: unusable when that section failed
      const uint32_t lfg = (job.index / f.ngx / 8) * f.nlfx + (job.index % f.ngx) / 8;
      st = f.status[lfg] == kOk ? DecodeAcGroup(br, f, job.index, nat, sc, pass) : (int) kErrBadStream;
    }
    if (st == kOk) st = DecodeModularGroup(br, f, job.index, sc, scratch.max_local_nodes, pass);
  }
  f.status[job.status_slot] = st;
}

// One warp per frame, after every entropy-coded section of the batch has been decoded: any section that failed (or was
// never reached) marks the frame bad.  The reconstruction kernels below skip bad frames -- their strategy / offset /
// multiplier planes are partly written garbage that must not drive loops or addresses.  (The host resolves the same
// statuses into the per-image error after the run, Batch::Finish.)
__global__ void __launch_bounds__(32) FrameStatusKernel(const FrameDev* frames, uint32_t nframes) {
  if (blockIdx.x >= nframes) return;
  const FrameDev& f = frames[blockIdx.x];
  int bad = 0;
  if (f.single_section) {
    if (threadIdx.x == 0) bad = f.status[f.num_lf_groups + f.num_groups] != kOk;
  } else {
    const uint32_t first = f.encoding == 0 ? 0u : f.num_lf_groups, last = f.num_lf_groups + f.num_groups;
    for (uint32_t i = first + threadIdx.x; i < last; i += 32) bad |= f.status[i] != kOk;
  }
  bad = __any_sync(0xFFFFFFFFu, bad);
  if (threadIdx.x == 0) *f.frame_bad = bad;
}

// ---- numeric ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) LfFinalKernel(const FrameDev f) {
  if (*f.frame_bad) return;
  const uint32_t cx = blockIdx.x * blockDim.x + threadIdx.x, cy = blockIdx.y;
  if (cx >= f.w8 || cy >= f.h8) return;
  if (cx == 0 && cy == 0) f.large_list[0] = 0;  // the inverse-transform kernel that follows appends to the list
  const LfMul m = MakeLfMul(f);
  float v[3];
  LfFinalCell(f, m, cx, cy, v);
  const size_t plane = (size_t) f.h8 * f.lf_stride, o = (size_t) cy * f.lf_stride + cx;
  f.lf[o] = v[0];
  f.lf[plane + o] = v[1];
  f.lf[2 * plane + o] = v[2];
}

constexpr int kReconThreads = 192;  // 3 channels x 64 columns/rows

__global__ void __launch_bounds__(kReconThreads, 4) ReconRegionKernel(const FrameDev f, const NumericTables* nt) {
  if (*f.frame_bad) return;
  extern __shared__ __align__(16) uint8_t smem[];
  RegionShared& sh = *reinterpret_cast<RegionShared*>(smem);
  ReconRegion(f, *nt, blockIdx.x, blockIdx.y, sh, (int) threadIdx.x, (int) blockDim.x, SyncThreads());
}

// One CTA per 64x64 region; it reconstructs the blocks whose top-left cell lies in the region but which are not
// contained in it (larger than 64 pixels, or straddling a region border) -- rare, so most CTAs exit after the scan.
// 128 threads x <= 128 registers: a CTA must fit beside the LF-group CTAs of another batch that overlap it (a 256-thread
// CTA at 254 registers needs a whole SM's register file and would stall its stream until an SM drains).
__global__ void __launch_bounds__(128, 4) ReconLargeKernel(const FrameDev f, const NumericTables* nt) {
  if (*f.frame_bad) return;
  __shared__ uint32_t todo[64];
  __shared__ uint32_t ntodo;
  if (threadIdx.x == 0) ntodo = 0;
  __syncthreads();
  if (threadIdx.x < 64) {
    const uint32_t bx = blockIdx.x * kRegionCells + (threadIdx.x & 7), by = blockIdx.y * kRegionCells + (threadIdx.x >> 3);
    if (bx < f.w8 && by < f.h8) {
      const uint8_t s = f.cell_strategy[(size_t) by * f.w8 + bx];
      if ((s & 0x80) && s != 0xFF && BlockNeedsLargePath(s & 0x7Fu, bx, by)) todo[atomicAdd(&ntodo, 1u)] = bx | (by << 16);
    }
  }
  __syncthreads();
  const uint32_t n = ntodo;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t e = todo[i];
    ReconLargeBlock(f, *nt, e & 0xFFFF, e >> 16, (int) threadIdx.x, (int) blockDim.x, SyncThreads());
    __syncthreads();
  }
}

// The same for the blocks ReconRegionTmaKernel listed (kernels_recon.cu): a fixed grid walks the list instead of 4096
// CTAs scanning for the rare block.
__global__ void __launch_bounds__(128, 4) ReconLargeListKernel(const FrameDev f, const NumericTables* nt) {
  if (*f.frame_bad) return;
  const uint32_t n = f.large_list[0];
  for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
    const uint32_t e = f.large_list[1 + i];
    ReconLargeBlock(f, *nt, e & 0xFFFF, e >> 16, (int) threadIdx.x, (int) blockDim.x, SyncThreads());
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) GaborishKernel(const FrameDev f, const float* src, float* dst) {
  if (*f.frame_bad) return;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= (int) f.width || y >= (int) f.height) return;
  StageGaborish(f, src, dst, x, y);
}

__global__ void __launch_bounds__(256) EpfKernel(const FrameDev f, const NumericTables* nt, int stage, const float* src, float* dst) {
  if (*f.frame_bad) return;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= (int) f.width || y >= (int) f.height) return;
  StageEpf(f, *nt, stage, src, dst, x, y);
}

__global__ void __launch_bounds__(256) ColorKernel(const FrameDev f, const ColorParams cp, const NumericTables* nt, const float* src,
                                                   OutputDesc out) {
  if (*f.frame_bad) return;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= (int) f.width || y >= (int) f.height) return;
  StageColorToRgba(f, cp, *nt, src, out, x, y);
}

__global__ void __launch_bounds__(256) Upsample2Kernel(const FrameDev f, const float* src, float* dst, uint32_t up_stride, uint32_t up_h) {
  if (*f.frame_bad) return;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= (int) f.width || y >= (int) f.height) return;
  StageUpsample2(f, src, dst, up_stride, up_h, x, y);
}

__global__ void __launch_bounds__(256) UpsampleAlpha2Kernel(const FrameDev f, const int32_t* src, uint32_t bits, int32_t* dst,
                                                             uint32_t up_stride) {
  if (*f.frame_bad) return;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= (int) f.width || y >= (int) f.height) return;
  StageUpsampleAlpha2(f, src, bits, dst, up_stride, x, y);
}

__global__ void __launch_bounds__(256) ModularToRgbaKernel(const FrameDev f, OutputDesc out) {
  if (*f.frame_bad) return;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= (int) f.width || y >= (int) f.height) return;
  if (!f.single_section && !f.global_serial) StageGlobalInverse(f, x, y);
  StageModularToRgba(f, out, x, y);
}

// Frame-level transforms that include a delta palette: one lane walks the whole image in raster order (modular.h:
// ApplyInverseTransforms).  Rare (libjxl's lossless encoder picks such palettes for about one file in a hundred).
__global__ void __launch_bounds__(32) GlobalInverseSerialKernel(const FrameDev f) {
  if (*f.frame_bad || threadIdx.x != 0) return;
  ModChannel planes[kMaxModPlanes];
  for (uint32_t c = 0; c < f.num_mod_channels && c < (uint32_t) kMaxModPlanes; ++c) {
    planes[c].data = f.mod + (size_t) c * f.height * f.mod_stride;
    planes[c].w = f.width;
    planes[c].h = f.height;
    planes[c].stride = f.mod_stride;
  }
  const int st = ApplyInverseTransforms(f.global_tr, f.global_nb_transforms, planes, f.meta, f.bit_depth);
  if (st != kOk) f.status[f.num_lf_groups] = st;
}

// Modular channels that the host decoded from the global stream of a small multi-section (progressive) frame: into their planes.
__global__ void __launch_bounds__(256) ScatterGlobalPlanesKernel(const FrameDev f) {
  if (*f.frame_bad) return;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= (int) f.width || y >= (int) f.height) return;
  for (uint32_t c = 0; c < f.num_coded; ++c)
    f.mod[(size_t) f.coded_plane[c] * f.height * f.mod_stride + (size_t) y * f.mod_stride + x] =
        f.global_planes[((size_t) c * f.height + y) * f.width + x];
}

// Frame-level transforms on the extra channels of a multi-section VarDCT frame (e.g. a palette on a lossless alpha).
__global__ void __launch_bounds__(256) ModularGlobalInverseKernel(const FrameDev f) {
  if (*f.frame_bad) return;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= (int) f.width || y >= (int) f.height) return;
  StageGlobalInverse(f, x, y);
}

__global__ void __launch_bounds__(256) PackKernel(const PackParams p) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= p.width || y >= p.height) return;
  PackPixel(p, x, y);
}

dim3 PixelGrid(uint32_t w, uint32_t h, uint32_t bx) { return dim3((w + bx - 1) / bx, h, 1); }

// Fills n16 16-byte words.  Used instead of cudaMemsetAsync for the coefficient planes: large memsets are executed by a
// copy engine, where they queue behind the 64 MiB result downloads of the other batches in flight.
__global__ void __launch_bounds__(256) FillKernel(uint4* p, size_t n16, uint32_t v) {
  const uint4 w = make_uint4(v, v, v, v);
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t) gridDim.x * blockDim.x) p[i] = w;
}

}  // namespace

void LaunchFill(void* p, size_t bytes, uint32_t value32, cudaStream_t stream) {
  if (!bytes) return;
  const size_t n16 = bytes / 16;  // callers pass 16-byte multiples (regions are 256-byte aligned)
  const unsigned blocks = (unsigned) std::min<size_t>((n16 + 255) / 256, 148 * 16);
  FillKernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<uint4*>(p), n16, value32);
  ++g_launches;
}

void LaunchPack(const PackParams& p, cudaStream_t stream) {
  PackKernel<<<PixelGrid(p.width, p.height, 256), 256, 0, stream>>>(p);
  ++g_launches;
}

extern std::atomic<uint64_t> g_launches_ac;
uint64_t KernelLaunchCount() { return g_launches.load() + g_launches_ac.load(); }

void LaunchSingleSectionFrames(const FrameDev* frames, const StreamJob* jobs, uint32_t njobs, NaturalOrders nat,
                               ScratchLayout scratch, cudaStream_t stream) {
  if (!njobs) return;
  SingleSectionKernel<<<njobs, kStreamBlockThreads, 0, stream>>>(frames, jobs, njobs, nat, scratch);
  ++g_launches;
}
void LaunchLfGroups(const FrameDev* frames, const StreamJob* jobs, uint32_t njobs, ScratchLayout scratch, cudaStream_t stream) {
  if (!njobs) return;
  // chains per CTA (JXLB_LF_WARPS, default 8 = two per scheduler); the CTA always asks for 200 KB of shared memory so that
  // no CTA of a dense kernel fits beside it
  static const int lf_warps = [] {
    const char* e = getenv("JXLB_LF_WARPS");
    const int v = e ? atoi(e) : 8;
    return v >= 1 && v <= kLfMaxWarps ? v : 8;
  }();
  static const bool exclusive = getenv("JXLB_LF_SHARED_SM") == nullptr;
  const int need = (int) (lf_warps * (kLfFastCodeBytes + kLfFastInts * 4));
  const int smem = exclusive ? std::max(need, 200 << 10) : need;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(LfGroupKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    // same (maximal) shared-memory carve-out for every long-running kernel: CTAs of kernels with different carve-outs
    // cannot share an SM, which would serialise the batches that overlap on different streams
    cudaFuncSetAttribute(LfGroupKernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  static const int warp_place = getenv("JXLB_SERIAL_PLACEMENT") == nullptr;
  LfGroupKernel<<<(njobs + lf_warps - 1) / lf_warps, lf_warps * 32, smem, stream>>>(frames, jobs, njobs, scratch, warp_place);
  ++g_launches;
}
void LaunchPassGroups(const FrameDev* frames, const StreamJob* jobs, uint32_t njobs, NaturalOrders nat, ScratchLayout scratch,
                      cudaStream_t stream) {
  if (!njobs) return;
  PassGroupKernel<<<njobs, kStreamBlockThreads, 0, stream>>>(frames, jobs, njobs, nat, scratch);
  ++g_launches;
}

void LaunchFrameStatus(const FrameDev* frames, uint32_t nframes, cudaStream_t stream) {
  if (!nframes) return;
  FrameStatusKernel<<<nframes, 32, 0, stream>>>(frames, nframes);
  ++g_launches;
}

void LaunchLfFinal(const FrameDev& f, cudaStream_t stream) {
  LfFinalKernel<<<PixelGrid(f.w8, f.h8, 256), 256, 0, stream>>>(f);
  ++g_launches;
}

void LaunchRecon(const FrameDev& f, const NumericTables* nt_dev, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(ReconRegionKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(RegionShared));
    // same (maximal) shared-memory carve-out for every long-running kernel: CTAs of kernels with different carve-outs
    // cannot share an SM, which would serialise the batches that overlap on different streams
    cudaFuncSetAttribute(ReconRegionKernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  dim3 grid((f.w8 + kRegionCells - 1) / kRegionCells, (f.h8 + kRegionCells - 1) / kRegionCells, 1);
  static bool large_configured = false;
  bool tma_done = false;
  if (!large_configured) {
    cudaFuncSetAttribute(ReconLargeKernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(ReconLargeListKernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(LfFinalKernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    large_configured = true;
  }
  // contained blocks: the persistent TMA kernel (kernels_recon.cu); the plain-load kernel above is its fallback
  tma_done = LaunchReconTma(f, nt_dev, stream);
  if (!tma_done) {
    ReconRegionKernel<<<grid, kReconThreads, sizeof(RegionShared), stream>>>(f, nt_dev);
    ++g_launches;
  }
  if (tma_done) ReconLargeListKernel<<<148 * 2, 128, 0, stream>>>(f, nt_dev);
  else ReconLargeKernel<<<grid, 128, 0, stream>>>(f, nt_dev);
  ++g_launches;
}

int LaunchFilters(const FrameDev& f, const NumericTables* nt_dev, cudaStream_t stream) {
  float* buf[2] = {f.xyb0, f.xyb1};
  int cur = 0;
  const dim3 grid = PixelGrid(f.width, f.height, 256);
  if (f.rf.gab) {
    GaborishKernel<<<grid, 256, 0, stream>>>(f, buf[cur], buf[cur ^ 1]);
    cur ^= 1;
    ++g_launches;
  }
  const int iters = f.rf.epf_iters;
  for (int stage = 0; stage < 3; ++stage) {
    const bool run = (stage == 0 && iters == 3) || (stage == 1 && iters >= 1) || (stage == 2 && iters >= 2);
    if (!run) continue;
    EpfKernel<<<grid, 256, 0, stream>>>(f, nt_dev, stage, buf[cur], buf[cur ^ 1]);
    cur ^= 1;
    ++g_launches;
  }
  return cur;
}

void LaunchColor(const FrameDev& f, const ColorParams& cp, const NumericTables* nt_dev, const float* src, OutputDesc out,
                 cudaStream_t stream) {
  ColorKernel<<<PixelGrid(f.width, f.height, 256), 256, 0, stream>>>(f, cp, nt_dev, src, out);
  ++g_launches;
}

void LaunchUpsample2(const FrameDev& f, const float* src, float* dst, uint32_t up_stride, uint32_t up_h, cudaStream_t stream) {
  Upsample2Kernel<<<PixelGrid(f.width, f.height, 256), 256, 0, stream>>>(f, src, dst, up_stride, up_h);
  ++g_launches;
}

void LaunchUpsampleAlpha2(const FrameDev& f, const int32_t* src, uint32_t bits, int32_t* dst, uint32_t up_stride,
                          cudaStream_t stream) {
  UpsampleAlpha2Kernel<<<PixelGrid(f.width, f.height, 256), 256, 0, stream>>>(f, src, bits, dst, up_stride);
  ++g_launches;
}

void LaunchScatterGlobalPlanes(const FrameDev& f, cudaStream_t stream) {
  ScatterGlobalPlanesKernel<<<PixelGrid(f.width, f.height, 256), 256, 0, stream>>>(f);
  ++g_launches;
}

void LaunchModularGlobalInverse(const FrameDev& f, cudaStream_t stream) {
  if (f.global_serial) GlobalInverseSerialKernel<<<1, 32, 0, stream>>>(f);
  else ModularGlobalInverseKernel<<<PixelGrid(f.width, f.height, 256), 256, 0, stream>>>(f);
  ++g_launches;
}

void LaunchModularToRgba(const FrameDev& f, OutputDesc out, cudaStream_t stream) {
  ModularToRgbaKernel<<<PixelGrid(f.width, f.height, 256), 256, 0, stream>>>(f, out);
  ++g_launches;
}

}  // namespace jxlb
