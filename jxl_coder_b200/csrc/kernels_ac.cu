// Lane-parallel AC coefficient decode (sm_100a).
//
// One PassGroup section is a serial ANS stream, so the parallelism is across sections: every LANE of a warp decodes
// its own 256x256 group, 32 groups per warp, all lanes of a CTA belonging to the same image so that the image's AC
// code (context map + alias tables, 16-100 KB) is staged once per CTA in shared memory and looked up with
// bank-parallel LDS instead of 32 scattered global loads.  To keep the 32 lanes converged the per-group decode is a
// flat state machine: every loop iteration decodes exactly one symbol (either a block's non-zero count or one
// coefficient), whatever block / channel / position the lane is at; block boundaries only touch predicated state.
// The per-group block lists the lanes iterate over are built by BuildGroupBlocksKernel from the LF-group placement.
//
// Functionally identical to DecodeAcGroup (vardct_sections.h), which stays as the reference implementation for the
// CPU tests and for streams this kernel does not take (LZ77-enabled AC codes, single-section frames).
#include <atomic>

#include "kernels.h"

namespace jxlb {

extern std::atomic<uint64_t> g_launches_ac;
std::atomic<uint64_t> g_launches_ac{0};

namespace {

constexpr int kAcCtaThreads = 128;
constexpr uint32_t kTopBytesPerLane = 96;  // 3 channels x 32 columns of the non-zero context row

__global__ void __launch_bounds__(128) BuildGroupBlocksKernel(const FrameDev* frames, const StreamJob* jobs, uint32_t njobs) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= njobs) return;
  const StreamJob job = jobs[j];
  const FrameDev& f = frames[job.frame];
  const uint32_t g = job.index;
  const uint32_t gx = g % f.ngx, gy = g / f.ngx;
  const uint32_t bx0 = gx * kGroupCells, by0 = gy * kGroupCells;
  const uint32_t bw = f.w8 - bx0 < kGroupCells ? f.w8 - bx0 : kGroupCells;
  const uint32_t bh = f.h8 - by0 < kGroupCells ? f.h8 - by0 : kGroupCells;
  uint32_t* out = f.group_blocks + (size_t) g * 1024;
  uint32_t n = 0;
  for (uint32_t by = 0; by < bh; ++by) {
    const uint8_t* srow = f.cell_strategy + (size_t) (by0 + by) * f.w8 + bx0;
    const uint16_t* qrow = f.cell_hfmul + (size_t) (by0 + by) * f.w8 + bx0;
    for (uint32_t bx = 0; bx < bw; ++bx) {
      const uint32_t sv = srow[bx];
      if (!(sv & 0x80) || sv == 0xFF) continue;
      out[n++] = bx | (by << 5) | ((sv & 0x7F) << 10) | (((uint32_t) qrow[bx] - 1u) << 16);
    }
  }
  f.group_nblocks[g] = n;
}

struct AcTables {
  uint8_t nnz[64];
  uint8_t freq[64];
};

__global__ void __launch_bounds__(kAcCtaThreads) AcLaneKernel(const FrameDev* frames, const AcCtaJob* jobs, NaturalOrders nat,
                                                              uint32_t smem_code_bytes) {
  extern __shared__ uint4 smem4[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(smem4);
  const AcCtaJob job = jobs[blockIdx.x];
  const FrameDev& f = frames[job.frame];
  const uint32_t tid = threadIdx.x;
  // ---- stage the image's AC code and the small context tables in shared memory
  const CodeHeader* gh = reinterpret_cast<const CodeHeader*>(f.ac_code);
  const uint32_t blob_bytes = gh->total_bytes;
  const bool in_smem = blob_bytes <= smem_code_bytes;
  if (in_smem) {
    const uint4* src = reinterpret_cast<const uint4*>(f.ac_code);
    for (uint32_t i = tid; i < blob_bytes / 16; i += kAcCtaThreads) smem4[i] = src[i];
  }
  AcTables* tabs = reinterpret_cast<AcTables*>(smem + smem_code_bytes);
  if (tid < 64) {
    tabs->nnz[tid] = (uint8_t) ZeroDensityNnzCtx(tid);
    tabs->freq[tid] = (uint8_t) ZeroDensityFreqCtx(tid);
  }
  uint8_t* top = smem + smem_code_bytes + sizeof(AcTables) + tid * kTopBytesPerLane;
  __syncthreads();
  if (tid >= job.ngroups) return;
  CodeView code;
  code.Bind(in_smem ? smem : f.ac_code);
  const uint32_t g = job.first_group + tid;
  const uint32_t gx = g % f.ngx, gy = g / f.ngx;
  const uint32_t gbx0 = gx * kGroupCells, gby0 = gy * kGroupCells;
  const uint32_t sec = 1 + f.num_lf_groups + 1 + g;
  BitReader br;
  br.Init(f.cs, f.cs_bytes, f.sec_bit_begin[sec], f.sec_bit_end[sec]);
  const uint32_t nbc = f.bctx.num_ctx;
  int status = kOk;
  const uint32_t hfp = br.Read(CeilLog2(f.num_hf_presets));
  if (hfp >= f.num_hf_presets) status = kErrBadStream;
  const uint32_t ctx_off = hfp * kContextsPerBlockCtx * nbc;
  SymbolReader sr;
  sr.Begin(code, br, nullptr, 0);
  const uint32_t* blocks = f.group_blocks + (size_t) g * 1024;
  const uint32_t nblocks = status == kOk ? f.group_nblocks[g] : 0;
  const bool has_lf_thr = f.bctx.num_lf_ctx > 1;
  const size_t cplane = (size_t) f.coef_h * f.coef_stride;
  // per-block state
  uint32_t bi = 0, ci = 0;
  uint32_t bx = 0, by = 0, t = 0, q = 1, cx = 1, covered = 1, l2 = 0, size = 64, ord = 0, kc_log = 3, lf_idx = 0;
  bool tall = true;
  // per-(block, channel) state
  bool in_coeffs = false;
  uint32_t nz = 0, k = 0, prev = 0, h0 = 0, c = 1;
  const uint16_t* order = nat.pool;
  int16_t* plane = f.coef;
  while (bi < nblocks) {
    uint32_t ctx;
    uint8_t* tp = top + c * 32;
    if (!in_coeffs) {
      if (ci == 0) {
        const uint32_t e = blocks[bi];
        bx = e & 31;
        by = (e >> 5) & 31;
        t = (e >> 10) & 63;
        q = (e >> 16) + 1;
        cx = StrategyCellsX(t);
        const uint32_t cy = StrategyCellsY(t);
        covered = cx * cy;
        l2 = (uint32_t) FloorLog2(covered);
        size = 64 * covered;
        ord = StrategyOrder(t);
        kc_log = 3 + (uint32_t) FloorLog2(cx > cy ? cx : cy);
        tall = cy >= cx;
        lf_idx = 0;
        if (has_lf_thr) {
          const size_t cell = (size_t) (gby0 + by) * f.lf_stride + gbx0 + bx;
          const size_t lplane = (size_t) f.h8 * f.lf_stride;
          const int32_t vy = f.lf_quant[cell], vx = f.lf_quant[lplane + cell], vb = f.lf_quant[2 * lplane + cell];
          uint32_t ix = 0, iy = 0, ib = 0;
          for (uint32_t i = 0; i < f.bctx.num_lf_thr[0]; ++i) ix += vx > f.bctx.lf_thr[0][i] ? 1 : 0;
          for (uint32_t i = 0; i < f.bctx.num_lf_thr[1]; ++i) iy += vy > f.bctx.lf_thr[1][i] ? 1 : 0;
          for (uint32_t i = 0; i < f.bctx.num_lf_thr[2]; ++i) ib += vb > f.bctx.lf_thr[2][i] ? 1 : 0;
          lf_idx = (ix * (f.bctx.num_lf_thr[2] + 1) + ib) * (f.bctx.num_lf_thr[1] + 1) + iy;
        }
      }
      c = ci == 0 ? 1u : ci == 1 ? 0u : 2u;  // coded Y, X, B
      tp = top + c * 32;
      uint32_t pred;
      if (bx == 0 && by == 0) pred = 32;
      else if (bx == 0) pred = tp[0];
      else if (by == 0) pred = tp[bx - 1];
      else pred = ((uint32_t) tp[bx] + tp[bx - 1] + 1u) >> 1;
      const uint32_t bc = BlockContext(f, ord, q, c, lf_idx);
      const uint32_t nzc = pred < 8 ? pred : (pred >= 64 ? 36 : 4 + pred / 2);
      ctx = ctx_off + nzc * nbc + bc;
      h0 = ctx_off + nbc * kNonZeroBuckets + kZeroDensityContexts * bc;
    } else {
      const uint32_t nl = (nz + covered - 1) >> l2;
      ctx = h0 + ((uint32_t) tabs->nnz[nl] + tabs->freq[k >> l2]) * 2 + prev;
    }
    const uint32_t u = ReadHybridUint(code, sr, br, ctx);
    bool advance = false;
    if (!in_coeffs) {
      nz = u;
      if (nz > size - covered) {
        status = kErrBadStream;
        break;
      }
      const uint8_t nzv = (uint8_t) ((nz + covered - 1) >> l2);
      for (uint32_t xx = 0; xx < cx; ++xx) tp[bx + xx] = nzv;
      if (nz == 0) {
        advance = true;
      } else {
        in_coeffs = true;
        k = covered;
        prev = nz > size / 16 ? 0 : 1;
        const uint32_t ooff = f.orders.offset[ord][c];
        order = (ooff & kOrderInFramePool) ? f.order_pool + (ooff & ~kOrderInFramePool) : nat.pool + ooff;
        plane = f.coef + c * cplane + (size_t) (gby0 + by) * 8 * f.coef_stride + (gbx0 + bx) * 8;
      }
    } else {
      if (u) {
        const int32_t v = UnpackSigned(u);
        if (v > 32767 || v < -32768) {
          status = kErrUnsupported;
          break;
        }
        const uint32_t pos = order[k];
        uint32_t r = pos >> kc_log, col = pos & ((1u << kc_log) - 1);
        if (tall) {
          const uint32_t tmp = r;
          r = col;
          col = tmp;
        }
        plane[(size_t) r * f.coef_stride + col] = (int16_t) v;
      }
      prev = u != 0 ? 1 : 0;
      nz -= prev;
      ++k;
      if (nz == 0) {
        in_coeffs = false;
        advance = true;
      } else if (k >= size) {
        status = kErrBadStream;
        break;
      }
    }
    if (advance) {
      if (++ci == 3) {
        ci = 0;
        ++bi;
      }
    }
  }
  if (status == kOk && !sr.FinalStateOk()) status = kErrBadStream;
  if (status == kOk && br.Overrun()) status = kErrTruncated;
  f.group_ac_end_bit[g] = br.Position();
  f.status[f.num_lf_groups + g] = status;
}

// Per-group modular data that follows the AC data in a PassGroup section (alpha / extra channels): one warp per group,
// lane 0 decodes, starting where the lane-parallel AC kernel stopped.
__global__ void __launch_bounds__(32) GroupModularKernel(const FrameDev* frames, const StreamJob* jobs, uint32_t njobs,
                                                         ScratchLayout scratch) {
  const uint32_t j = blockIdx.x;
  if (j >= njobs || threadIdx.x != 0) return;
  const StreamJob job = jobs[j];
  const FrameDev& f = frames[job.frame];
  if (f.status[job.status_slot] != kOk) return;
  uint8_t* p = scratch.base + (uint64_t) j * scratch.bytes_per_job;
  StreamScratch sc;
  sc.arena.Init(p, scratch.arena_bytes);
  p += (scratch.arena_bytes + 255u) & ~255u;
  sc.wp = reinterpret_cast<int32_t*>(p);
  sc.nzmap = nullptr;
  sc.lz77 = nullptr;
  sc.lz77_mask = 0;
  const uint32_t sec = 1 + f.num_lf_groups + 1 + job.index;
  BitReader br;
  br.Init(f.cs, f.cs_bytes, f.group_ac_end_bit[job.index], f.sec_bit_end[sec]);
  f.status[job.status_slot] = DecodeModularGroup(br, f, job.index, sc, scratch.max_local_nodes);
}

}  // namespace

void LaunchBuildGroupBlocks(const FrameDev* frames, const StreamJob* jobs, uint32_t njobs, cudaStream_t stream) {
  if (!njobs) return;
  BuildGroupBlocksKernel<<<(njobs + 127) / 128, 128, 0, stream>>>(frames, jobs, njobs);
  ++g_launches_ac;
}

uint32_t AcLaneSmemBytes(uint32_t code_bytes) {
  return ((code_bytes + 15u) & ~15u) + (uint32_t) sizeof(AcTables) + kAcCtaThreads * kTopBytesPerLane;
}

void LaunchAcLanes(const FrameDev* frames, const AcCtaJob* jobs, uint32_t njobs, NaturalOrders nat, uint32_t smem_code_bytes,
                   cudaStream_t stream) {
  if (!njobs) return;
  smem_code_bytes = (smem_code_bytes + 15u) & ~15u;
  const uint32_t smem = AcLaneSmemBytes(smem_code_bytes);
  static uint32_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(AcLaneKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    configured = smem;
  }
  AcLaneKernel<<<njobs, kAcCtaThreads, smem, stream>>>(frames, jobs, nat, smem_code_bytes);
  ++g_launches_ac;
}

void LaunchGroupModular(const FrameDev* frames, const StreamJob* jobs, uint32_t njobs, ScratchLayout scratch, cudaStream_t stream) {
  if (!njobs) return;
  GroupModularKernel<<<njobs, 32, 0, stream>>>(frames, jobs, njobs, scratch);
  ++g_launches_ac;
}

}  // namespace jxlb
