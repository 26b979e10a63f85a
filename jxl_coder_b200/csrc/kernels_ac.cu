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

constexpr int kAcMaxCtaGroups = 256;  // groups per CTA (all of the same image): 128, or 256 (JXLB_AC_GROUPS; see AcGroupsPerCta)
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
  if (f.status[(gy / 8) * f.nlfx + gx / 8] != kOk) {  // the LF group's placement is unusable: the group decodes nothing
    f.group_nblocks[g] = 0;
    return;
  }
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

// ZeroDensityFreqCtx (frame.h) in closed form for k < 64: 0, 0, 1 .. 14, then pairs 15 .. 22, then quads 23 .. 30.
// Saves a dependent shared-memory load on every coefficient.
__device__ __forceinline__ uint32_t FreqCtxArith(uint32_t k) {
  return k < 2 ? 0u : k < 16 ? k - 1u : k < 32 ? (k >> 1) + 7u : (k >> 2) + 15u;
}

// Bit reader of one lane: like BitReader, plus a one-word lookahead so that a refill never waits for memory (the load
// issued by refill n is consumed by refill n + 1).  Lanes refill at different iterations; keeping the load off the
// critical path is what stops one lane's L2 miss from stalling the other lanes of its warp.
struct LaneBits {
  const uint32_t* words;
  uint64_t buf;
  uint32_t nbits, widx, wend, nextw;
  __device__ __forceinline__ void From(const BitReader& br) {
    words = br.words;
    buf = br.buf;
    nbits = br.nbits;
    widx = br.widx;
    wend = br.wend;
    nextw = widx < wend ? __ldg(words + widx) : 0u;
  }
  __device__ __forceinline__ uint64_t Position() const { return (uint64_t) widx * 32 - nbits; }
  __device__ __forceinline__ void Refill() {
    if (nbits <= 32) {
      buf |= (uint64_t) nextw << nbits;
      nbits += 32;
      ++widx;
      nextw = widx < wend ? __ldg(words + widx) : 0u;
    }
  }
  __device__ __forceinline__ uint32_t Read(uint32_t n) {  // n <= 32
    Refill();
    const uint32_t v = (uint32_t) (buf & ((1ull << n) - 1));
    buf >>= n;
    nbits -= n;
    return v;
  }
};

// One hybrid integer of context `ctx` from an alias-table (ANS) code held in shared memory.
__device__ __forceinline__ uint32_t LaneReadUint(LaneBits& lb, uint32_t& state, const uint8_t* ctx_map, const uint2* alias,
                                                 const HybridCfg* cfgs, uint32_t log_alpha, uint32_t log_entry, uint32_t ctx) {
  const uint32_t cluster = ctx_map[ctx];
  const uint32_t res = state & (kAnsTabSize - 1);
  const uint32_t bucket = res >> log_entry;
  const uint32_t pos = res & ((1u << log_entry) - 1);
  const uint2 e = alias[(cluster << log_alpha) + bucket];  // {cutoff | right << 8 | freq0 << 16, offset1 | freq1 << 16}
  const HybridCfg cfg = cfgs[cluster];
  const bool hi = pos >= (e.x & 0xFFu);
  const uint32_t sym = hi ? ((e.x >> 8) & 0xFFu) : bucket;
  const uint32_t off = hi ? (e.y & 0xFFFFu) + pos : pos;
  const uint32_t freq = hi ? (e.y >> 16) : (e.x >> 16);
  uint32_t st = freq * (state >> kAnsTabBits) + off;
  if (st < (1u << 16)) {
    lb.Refill();
    st = (st << 16) | (uint32_t) (lb.buf & 0xFFFFu);
    lb.buf >>= 16;
    lb.nbits -= 16;
  }
  state = st;
  const uint32_t split = 1u << cfg.split_exp;
  if (sym < split) return sym;
  const uint32_t in_token = (uint32_t) cfg.msb + cfg.lsb;
  uint32_t nbits = cfg.split_exp - in_token + ((sym - split) >> in_token);
  if (nbits > 31) nbits = 31;
  const uint32_t low = sym & ((1u << cfg.lsb) - 1);
  const uint32_t tok = sym >> cfg.lsb;
  const uint32_t bits = lb.Read(nbits);
  return ((((1u << cfg.msb) | (tok & ((1u << cfg.msb) - 1))) << nbits | bits) << cfg.lsb) | low;
}

// kLanes = groups per warp (the other lanes idle): fewer groups per warp means more warps per SM to hide the latency
// of each lane's dependency chain, and fewer divergent paths to serialise inside a warp.
// kFast: alias-table code staged in shared memory (the common case); otherwise the general symbol reader is used.
template <int kLanes, bool kFast, int kAcCtaGroups>
__global__ void __launch_bounds__(kAcCtaGroups * 32 / kLanes) AcLaneKernel(const FrameDev* frames, const AcCtaJob* jobs, NaturalOrders nat,
                                                                           uint32_t smem_code_bytes) {
  constexpr int kThreads = kAcCtaGroups * 32 / kLanes;
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
    for (uint32_t i = tid; i < blob_bytes / 16; i += kThreads) smem4[i] = src[i];
  }
  AcTables* tabs = reinterpret_cast<AcTables*>(smem + smem_code_bytes);
  if (tid < 64) {
    tabs->nnz[tid] = (uint8_t) ZeroDensityNnzCtx(tid);
    tabs->freq[tid] = (uint8_t) ZeroDensityFreqCtx(tid);
  }
  __syncthreads();
  const uint32_t lane = tid & 31u, gi = (tid >> 5) * kLanes + lane;
  if (lane >= (uint32_t) kLanes || gi >= job.ngroups) return;
  uint8_t* top = smem + smem_code_bytes + sizeof(AcTables) + gi * kTopBytesPerLane;
  CodeView code;
  if (kFast) code.Bind(smem);  // the host only selects the fast kernel when every code of the batch fits (in_smem)
  else code.Bind(in_smem ? smem : f.ac_code);
  const uint32_t g = job.first_group + gi;
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
  // fast path state: prefetching bit reader + the code's tables as shared-memory pointers
  LaneBits lb;
  lb.From(br);
  uint32_t ans_state = sr.state;
  const uint8_t* s_ctx_map = code.ctx_map;
  const uint2* s_alias = reinterpret_cast<const uint2*>(code.alias);
  const HybridCfg* s_cfg = code.cfg;
  const uint32_t log_alpha = code.log_alpha, log_entry = code.log_entry;
  const uint32_t* blocks = f.group_blocks + (size_t) g * 1024;
  const uint32_t nblocks = status == kOk ? f.group_nblocks[g] : 0;
  const bool has_lf_thr = f.bctx.num_lf_ctx > 1;
  const uint32_t coef_stride = f.coef_stride;
  const size_t cplane = (size_t) f.coef_h * coef_stride;
  // per-block state
  uint32_t bi = 0, ci = 0;
  uint32_t bx = 0, by = 0, t = 0, q = 1, cx = 1, covered = 1, l2 = 0, size = 64, ord = 0, kc_log = 3, lf_idx = 0;
  bool tall = true;
  // per-(block, channel) state
  bool in_coeffs = false;
  uint32_t nz = 0, k = 0, prev = 0, h0 = 0, c = 1;
  uint32_t nnz_ctx = 0;  // tabs->nnz[(nz + covered - 1) >> l2], refreshed only when nz changes
  const uint16_t* order = nat.pool;
  int16_t* plane = f.coef;
  while (bi < nblocks) {
    uint32_t ctx;
    uint8_t* tp = top + c * 32;
    uint32_t pos_pref = 0;
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
      pos_pref = __ldg(order + k);  // used only if this coefficient is non-zero; issued early to hide its latency
      ctx = h0 + (nnz_ctx + FreqCtxArith(k >> l2)) * 2 + prev;
    }
    uint32_t u;
    if (kFast) {
      u = LaneReadUint(lb, ans_state, s_ctx_map, s_alias, s_cfg, log_alpha, log_entry, ctx);
    } else {
      u = ReadHybridUint(code, sr, br, ctx);
    }
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
        nnz_ctx = tabs->nnz[(nz + covered - 1) >> l2];
        const uint32_t ooff = f.orders.offset[ord][c];
        order = (ooff & kOrderInFramePool) ? f.order_pool + (ooff & ~kOrderInFramePool) : nat.pool + ooff;
        plane = f.coef + c * cplane + (size_t) (gby0 + by) * 8 * coef_stride + (gbx0 + bx) * 8;
      }
    } else {
      if (u) {
        const int32_t v = UnpackSigned(u);
        if (v > 32767 || v < -32768) {
          status = kErrUnsupported;
          break;
        }
        const uint32_t pos = pos_pref;
        uint32_t r = pos >> kc_log, col = pos & ((1u << kc_log) - 1);
        if (tall) {
          const uint32_t tmp = r;
          r = col;
          col = tmp;
        }
        plane[(size_t) r * coef_stride + col] = (int16_t) v;
      }
      prev = u != 0 ? 1 : 0;
      nz -= prev;
      if (prev) nnz_ctx = tabs->nnz[(nz + covered - 1) >> l2];
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
  uint64_t end_pos;
  bool overrun;
  if (kFast) {
    if (status == kOk && ans_state != (kAnsSignature << 16)) status = kErrBadStream;
    end_pos = lb.Position();
    overrun = end_pos > br.end_bit;
  } else {
    if (status == kOk && !sr.FinalStateOk()) status = kErrBadStream;
    end_pos = br.Position();
    overrun = br.Overrun();
  }
  if (status == kOk && overrun) status = kErrTruncated;
  f.group_ac_end_bit[g] = end_pos;
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
  sc.wp_ints = scratch.wp_ints;
  sc.nzmap = nullptr;
  sc.lz77 = nullptr;
  sc.lz77_mask = 0;
  if (job.lz_slot && scratch.lz_base) {
    sc.lz77 = reinterpret_cast<uint32_t*>(scratch.lz_base + (uint64_t) (job.lz_slot - 1) * scratch.lz_entries * 4u);
    sc.lz77_mask = scratch.lz_entries - 1;
  }
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
  return ((code_bytes + 15u) & ~15u) + (uint32_t) sizeof(AcTables) + kAcMaxCtaGroups * kTopBytesPerLane;
}

uint32_t AcGroupsPerCta() {
  // 256 groups per CTA (a whole 4096x4096 image) = 32 warps of 8 lanes on one SM: twice the warps to hide the lanes'
  // dependency chains, and a 64-image batch then holds 64 SMs instead of 128, leaving room for the LF stage and the
  // dense kernels of the other batches in flight.  The kernel needs 64 registers x 1024 threads = the whole register file.
  static const uint32_t v = [] {
    const char* e = getenv("JXLB_AC_GROUPS");
    const int n = e ? atoi(e) : 256;
    return (uint32_t) (n == 128 ? 128 : 256);
  }();
  return v;
}

namespace {
template <int kLanes, bool kFast, int kGroups>
void LaunchAcLanesT(const FrameDev* frames, const AcCtaJob* jobs, uint32_t njobs, NaturalOrders nat, uint32_t smem_code_bytes,
                    uint32_t smem, cudaStream_t stream) {
  static uint32_t configured = 0;
  if (smem > configured) {
    cudaFuncSetAttribute(AcLaneKernel<kLanes, kFast, kGroups>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    configured = smem;
  }
  AcLaneKernel<kLanes, kFast, kGroups><<<njobs, kGroups * 32 / kLanes, smem, stream>>>(frames, jobs, nat, smem_code_bytes);
}
int AcLanesPerWarp() {
  static int v = [] {
    const char* e = getenv("JXLB_AC_LANES");
    const int n = e ? atoi(e) : 8;
    return (n == 32 || n == 16 || n == 8) ? n : 8;
  }();
  return v;
}
}  // namespace

// fast: every job's AC code is an alias-table (ANS) code that fits the shared-memory budget.
void LaunchAcLanes(const FrameDev* frames, const AcCtaJob* jobs, uint32_t njobs, NaturalOrders nat, uint32_t smem_code_bytes,
                   bool fast, cudaStream_t stream) {
  if (!njobs) return;
  smem_code_bytes = (smem_code_bytes + 15u) & ~15u;
  const uint32_t smem = AcLaneSmemBytes(smem_code_bytes);
  const bool big = AcGroupsPerCta() == 256;
  if (!fast) {
    if (big) LaunchAcLanesT<32, false, 256>(frames, jobs, njobs, nat, smem_code_bytes, smem, stream);
    else LaunchAcLanesT<32, false, 128>(frames, jobs, njobs, nat, smem_code_bytes, smem, stream);
  } else if (big) {
    switch (AcLanesPerWarp()) {
      case 32: LaunchAcLanesT<32, true, 256>(frames, jobs, njobs, nat, smem_code_bytes, smem, stream); break;
      case 16: LaunchAcLanesT<16, true, 256>(frames, jobs, njobs, nat, smem_code_bytes, smem, stream); break;
      default: LaunchAcLanesT<8, true, 256>(frames, jobs, njobs, nat, smem_code_bytes, smem, stream); break;
    }
  } else {
    switch (AcLanesPerWarp()) {
      case 32: LaunchAcLanesT<32, true, 128>(frames, jobs, njobs, nat, smem_code_bytes, smem, stream); break;
      case 16: LaunchAcLanesT<16, true, 128>(frames, jobs, njobs, nat, smem_code_bytes, smem, stream); break;
      default: LaunchAcLanesT<8, true, 128>(frames, jobs, njobs, nat, smem_code_bytes, smem, stream); break;
    }
  }
  ++g_launches_ac;
}

void LaunchGroupModular(const FrameDev* frames, const StreamJob* jobs, uint32_t njobs, ScratchLayout scratch, cudaStream_t stream) {
  if (!njobs) return;
  GroupModularKernel<<<njobs, 32, 0, stream>>>(frames, jobs, njobs, scratch);
  ++g_launches_ac;
}

}  // namespace jxlb
