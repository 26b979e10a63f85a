// Host-side orchestration of one GPU: parse headers on the CPU, plan memory, upload the codestreams and the
// host-parsed tables, run the kernels of kernels.cu over the whole batch, pack to the requested Bitmap format and
// hand the results back (device memory or pinned host memory).
//
// Control flow per request mirrors decodeSampledImageImpl (/root/reference/jxlcoder/src/main/cpp/JniDecoding.cpp:45-331):
// checkDecodePreconditions -> decode (DecodeJpegXlOneShot semantics: RGBA u8, or u16 when bits_per_sample > 8 and
// api_level >= 26) -> [ICC] -> [rescale] -> [colour matrix when api_level < 34] -> ReformatColorConfig -> colour-space tag.
// There is no CPU fallback: every pixel is produced by a CUDA kernel, and a missing / failing device is an error.
#include "decoder.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <atomic>
#include <mutex>
#include <new>
#include <stdexcept>
#include <thread>

#include "color_matrix.h"
#include "color_params.h"
#include "frame_parser.h"
#include "kernels.h"
#include "numeric_tables.h"
#include "plan.h"
#include "resize.h"
#include "test_block.h"

namespace jxlb {

namespace {

#define CUDA_OK(expr)                                                                   \
  do {                                                                                  \
    cudaError_t e_ = (expr);                                                            \
    if (e_ != cudaSuccess) throw CudaError(std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

struct CudaError {
  std::string msg;
  explicit CudaError(std::string m) : msg(std::move(m)) {}
};

size_t Align256(size_t v) { return (v + 255) & ~(size_t) 255; }

// Runs fn(i) for i in [0, n) on up to hardware_concurrency host threads (header parsing and staging of a batch are
// independent per image).
// Host threads one call may use: the machine's cores divided among the processes of a multi-GPU job (one process per GPU
// under torchrun: LOCAL_WORLD_SIZE / WORLD_SIZE), so that 8 ranks do not run 8 x nproc parser threads on the same
// cores; JXLB_HOST_THREADS overrides.
size_t HostThreads() {
  static const size_t v = [] {
    if (const char* e = getenv("JXLB_HOST_THREADS")) return (size_t) std::max(1, atoi(e));
    unsigned hw = std::thread::hardware_concurrency();
    if (!hw) hw = 4;
    int world = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) world = std::max(1, atoi(e));
    else if (const char* e2 = getenv("WORLD_SIZE")) world = std::max(1, atoi(e2));
    return (size_t) std::max(2u, hw / (unsigned) world);
  }();
  return v;
}
// Host waits.  cudaEventSynchronize spins on a core; with several batches in flight per process and one process per GPU
// on a box with few cores per GPU (4 on the 8-GPU boxes) the spinning waiters starve the parse threads, while CUDA's
// blocking-sync waits wake up far too late here (measured: e2e 165 ms per step instead of 112).  JXLB_WAIT_MODE=poll
// polls the event every 50 us instead; the default ("auto") polls when the process has fewer than 6 host threads.
bool PollWaits() {
  static const bool v = [] {
    const char* e = getenv("JXLB_WAIT_MODE");
    if (e && !strcmp(e, "poll")) return true;
    if (e && !strcmp(e, "spin")) return false;
    return HostThreads() < 6;
  }();
  return v;
}
cudaError_t WaitEvent(cudaEvent_t ev) {
  if (!PollWaits()) return cudaEventSynchronize(ev);
  for (;;) {
    const cudaError_t r = cudaEventQuery(ev);
    if (r != cudaErrorNotReady) return r;
    std::this_thread::sleep_for(std::chrono::microseconds(50));
  }
}
cudaError_t WaitStream(cudaStream_t st) {
  if (!PollWaits()) return cudaStreamSynchronize(st);
  for (;;) {
    const cudaError_t r = cudaStreamQuery(st);
    if (r != cudaErrorNotReady) return r;
    std::this_thread::sleep_for(std::chrono::microseconds(50));
  }
}
template <class Fn>
void ParallelFor(size_t n, Fn fn) {
  size_t nt = std::min<size_t>(n, HostThreads());
  if (nt <= 1) {
    for (size_t i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<size_t> next{0};
  std::vector<std::thread> ths;
  auto work = [&]() {
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= n) return;
      fn(i);
    }
  };
  for (size_t t = 1; t < nt; ++t) ths.emplace_back(work);
  work();
  for (auto& t : ths) t.join();
}
constexpr size_t kMaxImageDeviceBytes = (size_t) 48 << 30;  // device planes of ONE image (a B200 has 180 GB)
constexpr uint32_t kMaxAcSmemCode = 196u << 10;  // AC code blobs up to this size are staged in shared memory (+ 24 KB of context rows per CTA)

struct DevBuffer {
  uint8_t* p = nullptr;
  size_t cap = 0;
  void Ensure(size_t n) {
    if (n <= cap) return;
    if (p) CUDA_OK(cudaFree(p));
    p = nullptr;
    cap = 0;
    size_t want = n + n / 8 + (1 << 20);
    CUDA_OK(cudaMalloc(&p, want));
    cap = want;
  }
};
struct PinnedBuffer {
  uint8_t* p = nullptr;
  size_t cap = 0;
  void Ensure(size_t n) {
    if (n <= cap) return;
    if (p) CUDA_OK(cudaFreeHost(p));
    p = nullptr;
    cap = 0;
    size_t want = n + n / 8 + (1 << 20);
    CUDA_OK(cudaMallocHost(&p, want));
    cap = want;
  }
};

// Pinned host result buffers are recycled: cudaMallocHost costs milliseconds per call.
struct HostPool {
  std::mutex mu;
  std::multimap<size_t, void*> free_list;
  std::map<void*, size_t> live;
  void* Get(size_t n) {
    std::lock_guard<std::mutex> l(mu);
    auto it = free_list.lower_bound(n);
    if (it != free_list.end() && it->first <= n + n / 4 + 4096) {
      void* p = it->second;
      live[p] = it->first;
      free_bytes -= it->first;
      free_list.erase(it);
      return p;
    }
    void* p = nullptr;
    if (cudaMallocHost(&p, n) != cudaSuccess) return nullptr;
    live[p] = n;
    return p;
  }
  // Freed buffers are kept for reuse up to a cap (JXLB_PINNED_POOL_MB, default 40 GiB: the idle sets of a few 64 x 4096^2 RGBA8 batches in flight);
  // beyond it the largest idle buffers go back to the driver, so a long-running process decoding varied sizes does not
  // accumulate page-locked memory without bound.
  size_t free_bytes = 0;
  static size_t Cap() {
    static const size_t v = [] {
      const char* e = getenv("JXLB_PINNED_POOL_MB");
      return (size_t) (e ? std::max(0, atoi(e)) : 40960) << 20;
    }();
    return v;
  }
  bool Put(void* p) {
    std::vector<void*> drop;
    {
      std::lock_guard<std::mutex> l(mu);
      auto it = live.find(p);
      if (it == live.end()) return false;
      free_list.emplace(it->second, p);
      free_bytes += it->second;
      live.erase(it);
      while (free_bytes > Cap() && !free_list.empty()) {
        auto big = std::prev(free_list.end());
        free_bytes -= big->first;
        drop.push_back(big->second);
        free_list.erase(big);
      }
    }
    for (void* d : drop) cudaFreeHost(d);
    return true;
  }
};
HostPool& Pool() {
  static HostPool* pool = new HostPool();
  return *pool;
}

struct BatchBuffers {
  DevBuffer const_buf, work_buf, scratch_buf, meta_buf, stage_out, final_out;
  DevBuffer xyb_ring;  // kXybRing XYB plane sets shared by the images of a batch (see Batch::Run)
  PinnedBuffer staging, status_host;
};

// One decode "slot": a pair of streams plus recycled batch buffers.  Concurrent jxlb_decode_batch calls (the reference's
// entry points are re-entrant and called from arbitrary app threads) take different slots, so the
// latency-bound LF stage of one call overlaps the throughput kernels and the downloads of another.
struct Slot {
  std::mutex mu;
  cudaStream_t stream = nullptr, copy_stream = nullptr, lf_stream = nullptr;
  BatchBuffers spare;
};
constexpr int kSlots = 8;

struct DeviceContext {
  int device = 0;
  std::mutex mu;  // guards lazily created state below
  Slot slots[kSlots];
  std::atomic<uint32_t> next_slot{0};
  // Stage gates.  The batches in flight on one GPU (concurrent jxlb_decode_batch calls, submitted batches, prepared
  // batches run asynchronously) would otherwise move in lockstep -- all in their LF stage together, then all in their
  // downloads together, with the PCIe link idle in between.  Each of the three stages of a run (LF sections / AC sections /
  // per-image reconstruction + downloads) admits a fixed number of batches at a time: a run enqueues, in front of the
  // stage, a wait for the end-of-stage event of the run that entered it `tokens` runs earlier.  The batches then spread
  // out into a software pipeline: while one reconstructs and downloads, the next decodes its AC sections and two more
  // run their LF chains (those hold 32 SMs each, see LfGroupKernel).  JXLB_STAGE_TOKENS="lf,ac,recon" (0 = no gate).
  struct StageGate {
    static constexpr int kRing = 64;
    cudaEvent_t ev[kRing]{};
    uint64_t count = 0;
    int tokens = 0;
    void Init(int t) {
      tokens = t;
      for (auto& e : ev) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    // Both are called with DeviceContext::mu held for the whole enqueue of a run, so the record of run k is always
    // enqueued before the wait of run k + tokens.
    // gated = false (prepared batches: everything resident in HBM, nothing to download) only takes a ticket, so that the
    // gated runs behind it still find its end-of-stage event
    uint64_t Enter(cudaStream_t s, bool gated) {
      const uint64_t k = count++;
      if (gated && tokens > 0 && k >= (uint64_t) tokens) CUDA_OK(cudaStreamWaitEvent(s, ev[(k - tokens) % kRing], 0));
      return k;
    }
    void Leave(uint64_t k, cudaStream_t s) { CUDA_OK(cudaEventRecord(ev[k % kRing], s)); }
  };
  StageGate gate_lf, gate_ac, gate_recon;
  // One stream carries every device-to-host copy of result pixels on this GPU: measured on B200 / PCIe Gen5 (tools/probes/gpu_pcie2.py),
  // 64 MiB copies into pinned buffers reach 48.5 GB/s from one or two streams and only 39 GB/s when four streams interleave.
  cudaStream_t d2h_stream = nullptr;
  cudaEvent_t origin = nullptr;  // JXLB_TIMELINE=1: stage boundaries of every run are printed relative to this event
  bool timeline = false;
  int lf_priority = 0;
  NumericTables* nt_dev = nullptr;
  NaturalOrders nat_dev{};
  bool ready = false;

  void Init(int dev) {
    device = dev;
    CUDA_OK(cudaSetDevice(dev));
    int prio_lo = 0, prio_hi = 0;
    CUDA_OK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    for (Slot& sl : slots) {
      CUDA_OK(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
      CUDA_OK(cudaStreamCreateWithFlags(&sl.copy_stream, cudaStreamNonBlocking));
      CUDA_OK(cudaStreamCreateWithPriority(&sl.lf_stream, cudaStreamNonBlocking, prio_hi));
    }
    lf_priority = prio_hi;
    CUDA_OK(cudaStreamCreateWithFlags(&d2h_stream, cudaStreamNonBlocking));
    {
      int t[3] = {2, 1, 1};
      if (const char* e = getenv("JXLB_STAGE_TOKENS")) sscanf(e, "%d,%d,%d", &t[0], &t[1], &t[2]);
      gate_lf.Init(t[0]);
      gate_ac.Init(t[1]);
      gate_recon.Init(t[2]);
    }
    timeline = getenv("JXLB_TIMELINE") != nullptr;
    CUDA_OK(cudaEventCreate(&origin));
    CUDA_OK(cudaEventRecord(origin, slots[0].stream));
    CUDA_OK(cudaEventSynchronize(origin));
    // constant tables
    const HostNumericTables& h = GetHostNumericTables();
    float* dq = nullptr;
    float* llf = nullptr;
    CUDA_OK(cudaMalloc(&dq, h.dequant_pool.size() * 4));
    CUDA_OK(cudaMalloc(&llf, h.llf_pool.size() * 4));
    CUDA_OK(cudaMemcpy(dq, h.dequant_pool.data(), h.dequant_pool.size() * 4, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(llf, h.llf_pool.data(), h.llf_pool.size() * 4, cudaMemcpyHostToDevice));
    NumericTables t = h.tables;
    t.dequant = dq;
    for (int l = 0; l < 6; ++l) t.llf[l] = llf + h.llf_off[l];
    CUDA_OK(cudaMalloc(&nt_dev, sizeof(NumericTables)));
    CUDA_OK(cudaMemcpy(nt_dev, &t, sizeof t, cudaMemcpyHostToDevice));
    const NaturalOrders& nh = NaturalOrderPoolHost();
    uint32_t total = nh.offset[kNumOrders - 1] + nh.size[kNumOrders - 1];
    uint16_t* pool = nullptr;
    CUDA_OK(cudaMalloc(&pool, total * 2));
    CUDA_OK(cudaMemcpy(pool, nh.pool, total * 2, cudaMemcpyHostToDevice));
    nat_dev = nh;
    nat_dev.pool = pool;
    ready = true;
  }
};

std::mutex g_ctx_mu;
std::map<int, std::unique_ptr<DeviceContext>> g_ctx;

DeviceContext* GetContext(int device) {
  std::lock_guard<std::mutex> l(g_ctx_mu);
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) throw CudaError("no CUDA device");
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) throw CudaError("no CUDA device available");
  if (device >= count) throw CudaError("CUDA device ordinal out of range");
  auto& c = g_ctx[device];
  if (!c) {
    c.reset(new DeviceContext());
    c->Init(device);
  }
  return c.get();
}

// Everything the host knows about one request after parsing.
struct Parsed {
  int status = JXLB_OK;
  std::string message;
  ByteVec cs;
  size_t cs_len = 0;
  ImageMetadata md;
  FrameHeader fh;
  FrameGlobals g;
  FramePlan plan;
  ColorParams cp{};
  // decode-stage output (what libjxl would hand to the reference)
  bool out16 = false;          // RGBA u16 (bits_per_sample > 8 and api_level >= 26)
  uint32_t depth = 8;          // bitDepth the reference tracks (8 or 16)
  bool has_alpha = false;      // hasAlphaInOrigin
  bool alpha_premultiplied = false;
  int format = JXLB_FORMAT_RGBA_8888;
  int color_space = JXLB_CS_NONE;
  // offsets into the batch buffers
  size_t const_off = 0, work_off = 0, stage_off = 0, stage_stride = 0;
  uint32_t status_base = 0;
  // rescale (decodeSampled with a target size): plan + offsets of its tables (const region) and images (work region)
  bool resize = false;
  ResizePlan rp;
  size_t rs_table_off = 0, rs_table_bytes = 0, rs_work_off = 0, rs_mid_bytes = 0, rs_scaled_bytes = 0;
  uint32_t out_w = 0, out_h = 0;   // size of the picture handed back
  // api_level < 34 colour pass (color_matrix.h): tables live in the const region
  // a cropped kReplace frame over an empty canvas (animations that store only each frame's bounding box): the frame is
  // decoded at its own size fw x fh and placed at (place_x0, place_y0) on a cleared canvas of the image size
  bool placed = false;
  int32_t place_x0 = 0, place_y0 = 0;
  uint32_t fw = 0, fh_ = 0;
  size_t place_off = 0;
  // orientation (applied right after the decode stage): oriented size and the offset of the oriented image (work region)
  uint32_t orient = 1, ow = 0, oh = 0;
  size_t or_off = 0;
  bool color_matrix = false;
  ColorMatrixPlan cmp;
  std::shared_ptr<ColorMatrixTables16> cmt16;  // sources deeper than 8 bits
  size_t cm_off = 0, cm_rows_off = 0;
  bool matrix_before_resize = false;  // getFrameImpl applies it before RescaleImage, decodeSampledImageImpl after
  // ---- frame composition (blending over earlier frames / reference slots) ----
  // One entry per frame of the codestream up to the target, as the frame headers describe them.
  struct SeenFrame {
    size_t begin_byte = 0, end_byte = 0;
    uint32_t frame_type = 0, fw = 0, fh = 0;
    int32_t x0 = 0, y0 = 0;
    bool covers_canvas = false, saves = false, save_before_ct = false, upsampled = false;
    uint32_t slot = 0;
    BlendingInfo blend, alpha_blend;
  };
  std::vector<SeenFrame> seen;   // filled for multi-frame codestreams
  size_t header_bytes = 0;       // image header (frames start byte-aligned after it)
  bool layer_only = false;       // hidden entry of a batch: this frame alone, decoded to straight RGBA at frame size
  bool composed = false;         // the picture is the target frame blended over `layers` (indices into Batch::ps, in frame order)
  std::vector<size_t> layers;
  std::vector<SeenFrame> layer_info;  // blend description of every layer, then of the target frame itself (layers.size() + 1 entries)
  size_t canvas_off = 0, canvas_bytes = 0;  // 5 canvases in the work region: reference slots 0 .. 3 and the current picture
  std::vector<int> chain;        // frames (indices into `seen`) the target needs, in order, set by ParseRequest
};

int Fail(Parsed* p, int status, const std::string& msg) {
  p->status = status;
  p->message = msg;
  return status;
}

// checkDecodePreconditions (Support.cpp:35-92)
int CheckPreconditions(const jxlb_request& r, int api, Parsed* p) {
  if (r.color_config < 1 || r.color_config > 6)
    return Fail(p, JXLB_BAD_ARG, "Invalid Color Config: " + std::to_string(r.color_config) + " was passed");
  if (r.color_config == JXLB_CONFIG_RGBA_1010102 && api < 33)
    return Fail(p, JXLB_BAD_ARG, "Color Config RGBA_1010102 supported only 33+ OS version but current is: " + std::to_string(api));
  if (r.color_config == JXLB_CONFIG_RGBA_F16 && api < 26)
    return Fail(p, JXLB_BAD_ARG, "Color Config RGBA_1010102 supported only 26+ OS version but current is: " + std::to_string(api));
  if (r.color_config == JXLB_CONFIG_HARDWARE && api < 29)
    return Fail(p, JXLB_BAD_ARG, "Color Config HARDWARE supported only 29+ OS version but current is: " + std::to_string(api));
  if (r.scale_mode < 1 || r.scale_mode > 3) return Fail(p, JXLB_BAD_ARG, "Invalid Scale Mode was passed");
  if (r.filter < 1 || r.filter > 10) return Fail(p, JXLB_BAD_ARG, "Invalid Sampler: " + std::to_string(r.filter) + " was passed");
  return JXLB_OK;
}

int MapParse(int st) { return st == kParseUnsupported ? JXLB_UNSUPPORTED : JXLB_INVALID_JXL; }

int ColorSpaceTag(const ImageMetadata& md, int api) {
  if (api < 34) return JXLB_CS_NONE;
  // JniDecoding.cpp:236-253
  const uint32_t prim = md.color.primaries, tf = md.color.have_gamma ? 0xFFFFu : md.color.transfer;
  if (prim == 9 && tf == 16) return JXLB_CS_BT2020_PQ;
  if (prim == 9 && tf == 18) return JXLB_CS_BT2020_HLG;
  if (prim == 11 && tf == 13) return JXLB_CS_DISPLAY_P3;
  if (prim == 1 && tf == 8) return JXLB_CS_LINEAR_SRGB;
  if (prim == 11 && tf == 17) return JXLB_CS_DCI_P3;
  if (prim == 1 && tf == 1) return JXLB_CS_BT2020_HLG;  // sic: the reference tags sRGB primaries + 709 TF as Hlg2100
  return JXLB_CS_SRGB;
}

// layer_only: decode displayed-or-not frame number `target_frame` (counting EVERY frame of the codestream) alone, as a
// layer of a composition: no independence check, no placement / orientation / rescale / colour pass / reformat.
void ParseRequest(const jxlb_request& r, int api, Parsed* p, int target_frame = -1, bool layer_only = false) {
  p->layer_only = layer_only;
  if (CheckPreconditions(r, api, p)) return;
  if (!r.data || r.len == 0) {
    Fail(p, JXLB_INVALID_JXL, "empty input");
    return;
  }
  int st = ExtractCodestream(r.data, r.len, &p->cs, &p->cs_len);
  if (st) {
    Fail(p, JXLB_INVALID_JXL, "not a JPEG XL file");
    return;
  }
  std::string err;
  uint64_t frame_bit = 0;
  st = ParseImageHeader(p->cs.data(), p->cs.size(), p->cs_len, &p->md, &frame_bit, &err);
  if (st) {
    Fail(p, MapParse(st), err);
    return;
  }
  const ImageMetadata& md = p->md;
  // DecodeJpegXlOneShot: 16-bit output iff bits_per_sample > 8 && allowedFloats (api >= 26)
  p->out16 = md.bits_per_sample > 8 && api >= 26;
  p->depth = p->out16 ? 16 : 8;
  {
    const uint64_t bytes = (uint64_t) md.xsize * md.ysize * 4 * (p->out16 ? 2 : 1);
    if (bytes >= 0x7FFFFFFFull) {
      Fail(p, JXLB_INVALID_SIZE, "Invalid image size exceed allowance, current size w: " + std::to_string(md.xsize) + ", h: " + std::to_string(md.ysize));
      return;
    }
  }
  p->has_alpha = false;
  p->alpha_premultiplied = false;
  if (!md.extra.empty()) {
    int ai = md.alpha_channel();
    if (ai >= 0 && md.extra[ai].bits > 0) {
      p->has_alpha = true;
      p->alpha_premultiplied = md.extra[ai].alpha_premultiplied;
    }
  }
  // libjxl applies the codestream orientation to the pixels it hands out (keep_orientation is never set by
  // DecodeJpegXlOneShot): the decoded picture is ow x oh
  p->orient = md.orientation;
  p->ow = md.orientation >= 5 ? md.ysize : md.xsize;
  p->oh = md.orientation >= 5 ? md.xsize : md.ysize;
  if (md.float_samples) {
    Fail(p, JXLB_UNSUPPORTED, "float samples");
    return;
  }
  // frames: decode the last one (or displayed frame `target_frame`); earlier frames must not be needed
  int nframes = 0, displayed = 0;
  p->header_bytes = (size_t) (frame_bit / 8);
  bool saved[4] = {false, false, false, false};  // reference slots written by the frames before the target
  for (;;) {
    p->fh = FrameHeader();
    st = ParseFrameHeader(p->cs.data(), p->cs.size(), p->cs_len, md, frame_bit, &p->fh, &err);
    if (st) {
      Fail(p, MapParse(st), err);
      return;
    }
    ++nframes;
    const bool shown = p->fh.frame_type == 0 || p->fh.frame_type == 3;
    if (shown) ++displayed;
    {
      Parsed::SeenFrame sf;
      sf.begin_byte = (size_t) (frame_bit / 8);
      sf.end_byte = (size_t) p->fh.end_byte;
      sf.frame_type = p->fh.frame_type;
      sf.fw = p->fh.coded_w;
      sf.fh = p->fh.coded_h;
      sf.x0 = p->fh.have_crop ? p->fh.x0 : 0;
      sf.y0 = p->fh.have_crop ? p->fh.y0 : 0;
      sf.covers_canvas = !p->fh.have_crop && p->fh.coded_w == md.xsize && p->fh.coded_h == md.ysize;
      sf.saves = !p->fh.is_last && (p->fh.frame_type == 2 || (p->fh.frame_type != 1 && (p->fh.duration == 0 || p->fh.save_as_reference != 0)));
      sf.slot = p->fh.save_as_reference & 3;
      sf.save_before_ct = p->fh.save_before_ct;
      sf.upsampled = p->fh.upsampling != 1;
      sf.blend = p->fh.blend;
      const int ai = md.alpha_channel();
      sf.alpha_blend = (ai >= 0 && (size_t) ai < p->fh.ec_blend.size()) ? p->fh.ec_blend[ai] : p->fh.blend;
      p->seen.push_back(sf);
    }
    if (layer_only ? nframes - 1 == target_frame : (target_frame >= 0 ? (shown && displayed - 1 == target_frame) : p->fh.is_last)) break;
    if (p->fh.is_last) {
      Fail(p, JXLB_BAD_ARG, "frame index out of range");
      return;
    }
    if (p->fh.frame_type == 2 || (p->fh.frame_type != 1 && (p->fh.duration == 0 || p->fh.save_as_reference != 0)))
      saved[p->fh.save_as_reference & 3] = true;
    frame_bit = p->fh.end_byte * 8;
  }
  const FrameHeader& fh = p->fh;
  // the picture a frame yields: its coded size, or the frame's size in image pixels when it is coded at half resolution
  p->fw = fh.upsampling == 2 ? fh.width : fh.coded_w;
  p->fh_ = fh.upsampling == 2 ? fh.height : fh.coded_h;
  const bool covers_canvas = !fh.have_crop && p->fw == md.xsize && p->fh_ == md.ysize;
  if (!layer_only && (nframes > 1 || !covers_canvas)) {
    // Frames are decoded as independent pictures.  That is what a kReplace frame is when it covers the canvas, or when
    // the canvas it is laid over is empty (its source slots were never written): then the picture is the frame at its
    // crop position over cleared pixels.  Anything else needs the earlier frames (composition): refused.
    bool independent = fh.blend.mode == 0;
    for (const BlendingInfo& b : fh.ec_blend) independent = independent && b.mode == 0;
    if (!covers_canvas) {
      independent = independent && !saved[fh.blend.source & 3];
      for (const BlendingInfo& b : fh.ec_blend) independent = independent && !saved[b.source & 3];
    }
    if (!independent) {
      // Composition: the picture is this frame blended over what earlier frames left in the reference slots.  Walk back
      // from the target to find the frames it really depends on (a frame matters when it is stored in a slot that a later
      // needed frame reads as its background).
      const std::vector<Parsed::SeenFrame>& F = p->seen;
      const int T = (int) F.size() - 1;
      auto needs_bg = [&](const Parsed::SeenFrame& f) {
        return !(f.covers_canvas && f.blend.mode == 0 && (!p->has_alpha || f.alpha_blend.mode == 0));
      };
      bool ok = true;
      std::string why;
      uint32_t need = 0;  // bit s: the content of reference slot s (as of the frame being looked at) is needed
      std::vector<int> chain;
      auto check = [&](const Parsed::SeenFrame& f) {
        if (f.frame_type == 1) ok = false, why = "LF frame";
        else if (f.frame_type == 2 && !f.covers_canvas) ok = false, why = "reference-only frame that is not canvas-sized (patch source)";
        else if (f.save_before_ct && f.saves) ok = false, why = "reference frame saved before the colour transform";
        else if (f.upsampled) ok = false, why = "upsampled frame in a composition";
        else if (p->has_alpha && f.alpha_blend.source != f.blend.source && needs_bg(f)) ok = false, why = "alpha blended from a different reference slot than colour";
        else if (f.blend.mode > 4 || f.alpha_blend.mode > 4) ok = false, why = "blend mode";
      };
      check(F[T]);
      if (needs_bg(F[T])) need |= 1u << (F[T].blend.source & 3);
      for (int k = T - 1; k >= 0 && ok && need; --k) {
        if (!F[k].saves || !(need & (1u << F[k].slot))) continue;
        check(F[k]);
        chain.push_back(k);
        need &= ~(1u << F[k].slot);
        if (needs_bg(F[k])) need |= 1u << (F[k].blend.source & 3);
      }
      if (!ok) {
        Fail(p, JXLB_UNSUPPORTED, "multi-frame image needing composition: " + why);
        return;
      }
      std::reverse(chain.begin(), chain.end());
      p->composed = true;
      p->chain = chain;  // what is still in `need` was never written: an empty canvas
    }
  }
  if (!covers_canvas && !p->composed && !layer_only) {
    if (md.orientation != 1 || fh.upsampling != 1) {
      Fail(p, JXLB_UNSUPPORTED, "cropped frame with orientation or upsampling");
      return;
    }
    p->placed = true;
    p->place_x0 = fh.have_crop ? fh.x0 : 0;
    p->place_y0 = fh.have_crop ? fh.y0 : 0;
  }
  st = ParseFrameGlobals(p->cs.data(), p->cs.size(), md, fh, &p->g, &err);
  if (st) {
    Fail(p, MapParse(st), err);
    return;
  }
  if (fh.encoding == 0) {
    if (!md.xyb_encoded) {
      Fail(p, JXLB_UNSUPPORTED, "VarDCT frame that is not XYB encoded");
      return;
    }
    st = MakeColorParams(md, &p->cp, &err);
    if (st) {
      Fail(p, MapParse(st), err);
      return;
    }
  } else if (md.xyb_encoded) {
    Fail(p, JXLB_UNSUPPORTED, "XYB-encoded modular frame");
    return;
  } else if (fh.rf.gab || fh.rf.epf_iters) {
    // libjxl runs Gaborish / EPF on modular frames too when the frame header asks for them (its encoder never does for
    // lossless); the modular path here has no filter stage
    Fail(p, JXLB_UNSUPPORTED, "restoration filters on a modular frame");
    return;
  }
  for (const ExtraChannelInfo& ec : md.extra)
    if (ec.bits > 16 || ec.is_float) {
      Fail(p, JXLB_UNSUPPORTED, "extra channel sample type");
      return;
    }
  if (md.extra.size() > 4) {
    Fail(p, JXLB_UNSUPPORTED, "more than 4 extra channels");
    return;
  }
  // DecodeJpegXlOneShot keeps the enum colour encoding only for these transfer functions (interop/JxlDecoding.cpp:125-133,
  // operator precedence as written there); every other image gets libjxl's synthesised ICC profile and is converted to
  // sRGB by lcms2 (convertUseDefinedColorSpace, JniDecoding.cpp:104-114) -- not restated, so refused rather than handed
  // back unconverted
  {
    const uint32_t tf = md.color.have_gamma ? 0xFFFFu : md.color.transfer;
    const bool prefer = (md.color.color_space == 0 && tf == 18) || tf == 16 || tf == 17 || tf == 1 || tf == 13 || tf == 0xFFFFu;
    if (!prefer) {
      Fail(p, JXLB_UNSUPPORTED, "colour encoding that the reference converts through its ICC path (lcms2)");
      return;
    }
  }
  if (layer_only) {  // a layer of a composition: the frame's own samples, nothing else
    p->orient = 1;
    p->out_w = p->fw;
    p->out_h = p->fh_;
    p->format = JXLB_FORMAT_RGBA_8888;
    p->color_space = JXLB_CS_NONE;
    MakeFramePlan(md, fh, p->g, p->cs.size(), &p->plan);
    return;
  }
  // rescale (JniDecoding.cpp:116-136)
  const bool use_sampler = (r.width > 0 || r.height > 0) && (r.width != 0 && r.height != 0);
  p->out_w = p->ow;
  p->out_h = p->oh;
  if (use_sampler) {
    // RescaleImage (SizeScaler.cpp:38-144) -> weave_scale_u8; the u16 path (f32 arithmetic) is not pinned and is refused
    if (p->out16) {
      Fail(p, JXLB_UNSUPPORTED, "rescale of 16-bit sources");
      return;
    }
    const int rs = MakeResizePlan(p->ow, p->oh, r.width, r.height, r.scale_mode, r.filter, p->has_alpha, &p->rp);
    if (rs == kResizeUnsupported) {
      Fail(p, JXLB_UNSUPPORTED, "rescale configuration");
      return;
    }
    if (rs != kResizeOk) {
      Fail(p, JXLB_INVALID_SIZE, "invalid target size");
      return;
    }
    p->resize = true;
    p->out_w = p->rp.out_w;
    p->out_h = p->rp.out_h;
  }
  if (api < 34) {  // JniDecoding.cpp:138-228
    bool needed = false;
    if (p->out16) p->cmt16.reset(new ColorMatrixTables16());
    if (!MakeColorMatrixPlan(md, &needed, &p->cmp, p->cmt16.get())) {
      Fail(p, JXLB_UNSUPPORTED, "api_level < 34 colour pass for this colour encoding");
      return;
    }
    if (!needed) p->cmt16.reset();
    p->color_matrix = needed;
    p->matrix_before_resize = target_frame >= 0;
  }
  // ReformatColorConfig: resolve Default (ReformatBitmap.cpp:52-63)
  int cfg = r.color_config;
  if (cfg == JXLB_CONFIG_DEFAULT) {
    if (p->depth > 8 && api >= 26) cfg = (api >= 33 && !p->has_alpha) ? JXLB_CONFIG_RGBA_1010102 : JXLB_CONFIG_RGBA_F16;
    else cfg = JXLB_CONFIG_RGBA_8888;
  }
  if (cfg == JXLB_CONFIG_HARDWARE) {
    Fail(p, JXLB_ERROR, "Error while decoding: Cannot load hardware buffers API");
    return;
  }
  p->format = cfg == JXLB_CONFIG_RGBA_8888 ? JXLB_FORMAT_RGBA_8888 : cfg == JXLB_CONFIG_RGBA_F16 ? JXLB_FORMAT_RGBA_F16
              : cfg == JXLB_CONFIG_RGB_565 ? JXLB_FORMAT_RGB_565 : JXLB_FORMAT_RGBA_1010102;
  p->color_space = ColorSpaceTag(md, api);
  MakeFramePlan(md, fh, p->g, p->cs.size(), &p->plan);
}

}  // namespace

// ---- batch object ---------------------------------------------------------------------------------------------------
struct Batch {
  DeviceContext* ctx = nullptr;
  int api_level = 34;
  size_t n = 0;
  std::vector<Parsed> ps;
  std::vector<uint32_t> frame_of;
  std::vector<FrameDev> frames;
  std::vector<StreamJob> jobs_single, jobs_lf, jobs_groups;   // jobs_groups: groups decoded by the one-warp-per-section kernel
  std::vector<StreamJob> jobs_lane_groups, jobs_lane_mod;     // groups taken by the lane-parallel AC kernel (+ their modular tails)
  uint32_t lz_slots = 0, lz_entries = 0;                      // LZ77 windows handed to modular streams (StreamJob::lz_slot)
  std::vector<AcCtaJob> jobs_ac_cta;
  uint32_t ac_smem_code_bytes = 0;
  bool ac_fast = true;        // every lane-decoded image has an alias-table (ANS) AC code
  ScratchLayout sl_single{}, sl_lf{}, sl_grp{};
  size_t const_total = 0, work_total = 0, stage_total = 0, final_total = 0, meta_total = 0;
  // Reconstruction runs image by image (LF final -> inverse transforms -> filters + colour + pack -> download), so the
  // 12 B/pixel XYB planes between the two kernels only ever exist for a few images at a time: a ring of kXybRing plane
  // sets replaces one set per image (12 GB for a 64 x 4096^2 batch), and a finished image starts its download while the
  // next ones are still being reconstructed.
  static constexpr size_t kXybRing = 4;
  size_t xyb_slot_bytes = 0;
  // per-kernel timing: every kTimeEvery-th VarDCT image of a run brackets its inverse-transform and filter kernels with
  // events (stage times = mean over the sampled launches x number of images)
  static constexpr size_t kTimeEvery = 4;
  size_t n_vardct = 0;
  uint32_t nframes = 0, status_total = 0;
  std::vector<size_t> final_off, final_bytes;
  const FrameDev* frames_d = nullptr;
  const StreamJob* jobs_single_d = nullptr;
  const StreamJob* jobs_lf_d = nullptr;
  const StreamJob* jobs_groups_d = nullptr;
  const StreamJob* jobs_lane_groups_d = nullptr;
  const StreamJob* jobs_lane_mod_d = nullptr;
  const AcCtaJob* jobs_ac_cta_d = nullptr;
  BatchBuffers own;           // buffers owned by this batch
  BatchBuffers* buf = nullptr;
  // streams: a slot's (one-shot decodes) or the batch's own (prepared batches, so that two prepared batches overlap)
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  // The LF stage runs on a stream of its own at the highest priority: its CTAs (one SM each, see LfGroupKernel) are placed
  // as soon as an SM drains instead of queueing behind the dense kernels of the batches in flight.
  cudaStream_t lf_stream = nullptr;
  cudaEvent_t ev_uploaded = nullptr, ev_lf_done = nullptr;
  bool own_streams = false;
  std::mutex mu;              // serialises Run / Finish / Fetch on this batch
  // Event sets: ev = the set of the current run.  Asynchronous runs (RunBatch(sync = false)) rotate through kEventSets
  // sets so that the stage times of every run can still be read after the final wait.
  static constexpr int kEventSets = 16;
  cudaEvent_t ev_ring[kEventSets][10]{};
  cudaEvent_t* ev = ev_ring[0];
  int pending_runs = 0, run_index = 0;
  double stage_sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int stage_runs = 0;
  std::vector<cudaEvent_t> sample_ev[kEventSets];  // 3 events per sampled image: before IDCT, after IDCT, after filters
  size_t sampled_in_run[kEventSets] = {};
  void EnsureSampleEvents() {
    const size_t want = 3 * ((n_vardct + kTimeEvery - 1) / kTimeEvery + 1);
    std::vector<cudaEvent_t>& v = sample_ev[run_index];
    while (v.size() < want) {
      cudaEvent_t e = nullptr;
      CUDA_OK(cudaEventCreate(&e));
      v.push_back(e);
    }
  }
  cudaEvent_t span_start = nullptr, span_end = nullptr;  // first run start / last run end since ResetStats
  bool span_armed = true;
  bool events = false, upload_timed = false;
  bool finished[kEventSets] = {};
  // e2e path: pinned host destinations; image i is copied out on the copy stream as soon as its last kernel is done,
  // so the download of image i overlaps the reconstruction of image i + 1
  std::vector<void*> host_dst;
  std::vector<cudaEvent_t> img_ev;
  bool uploaded = false, ran = false;
  BatchTimings tm;

  ~Batch() {
    if (events)
      for (auto& set : ev_ring)
        for (auto& e : set) cudaEventDestroy(e);
    for (auto& v : sample_ev)
      for (auto& e : v) cudaEventDestroy(e);
    if (span_start) cudaEventDestroy(span_start);
    if (span_end) cudaEventDestroy(span_end);
    if (own_streams) {
      if (stream) cudaStreamDestroy(stream);
      if (copy_stream) cudaStreamDestroy(copy_stream);
      if (lf_stream) cudaStreamDestroy(lf_stream);
    }
    if (ev_uploaded) cudaEventDestroy(ev_uploaded);
    if (ev_lf_done) cudaEventDestroy(ev_lf_done);
    for (auto& e : img_ev)
      if (e) cudaEventDestroy(e);
    for (void* h : host_dst)
      if (h) Pool().Put(h);
    auto freed = [](DevBuffer& b) {
      if (b.p) cudaFree(b.p);
      b.p = nullptr;
    };
    auto freeh = [](PinnedBuffer& b) {
      if (b.p) cudaFreeHost(b.p);
      b.p = nullptr;
    };
    freed(own.const_buf);
    freed(own.work_buf);
    freed(own.scratch_buf);
    freed(own.meta_buf);
    freed(own.stage_out);
    freed(own.final_out);
    freed(own.xyb_ring);
    freeh(own.staging);
    freeh(own.status_host);
  }

  // Requests whose picture is a composition (ParseRequest set `composed` and the chain of frames the target depends on):
  // every frame of the chain becomes a hidden entry of the batch -- a mini codestream (image header + that frame's bytes)
  // decoded alone to straight RGBA at frame size -- and the owner composites them in order in Run().
  std::vector<std::vector<uint8_t>> layer_streams;  // owns the mini codestreams until they have been parsed
  void ExpandCompositions(const jxlb_request* reqs, const int32_t* frame_index) {
    (void) frame_index;
    const size_t nreq = ps.size();
    struct Todo {
      size_t owner;
      int frame;
    };
    std::vector<Todo> todo;
    for (size_t i = 0; i < nreq; ++i) {
      Parsed& p = ps[i];
      if (p.status != JXLB_OK || !p.composed) continue;
      for (int k : p.chain) todo.push_back(Todo{i, k});
    }
    if (todo.empty()) return;
    const size_t first_hidden = ps.size();
    ps.resize(first_hidden + todo.size());
    layer_streams.resize(todo.size());
    ParallelFor(todo.size(), [&](size_t t) {
      const Parsed& owner = ps[todo[t].owner];
      const Parsed::SeenFrame& sf = owner.seen[todo[t].frame];
      std::vector<uint8_t>& mini = layer_streams[t];
      mini.assign(owner.cs.begin(), owner.cs.begin() + owner.header_bytes);
      mini.insert(mini.end(), owner.cs.begin() + sf.begin_byte, owner.cs.begin() + sf.end_byte);
      jxlb_request r = reqs[todo[t].owner];
      r.data = mini.data();
      r.len = mini.size();
      r.width = r.height = -1;
      r.color_config = JXLB_CONFIG_RGBA_8888;
      Parsed& hp = ps[first_hidden + t];
      try {
        ParseRequest(r, api_level, &hp, 0, /*layer_only=*/true);
      } catch (const std::exception&) {
        hp = Parsed();
        Fail(&hp, JXLB_OOM, "Not enough memory to decode this image");
      }
      hp.layer_only = true;
    });
    layer_streams.clear();
    for (size_t t = 0; t < todo.size(); ++t) {
      Parsed& owner = ps[todo[t].owner];
      const Parsed& hp = ps[first_hidden + t];
      if (hp.status != JXLB_OK && owner.status == JXLB_OK) {
        Fail(&owner, hp.status, hp.message);
        continue;
      }
      owner.layers.push_back(first_hidden + t);
      owner.layer_info.push_back(owner.seen[todo[t].frame]);
    }
    for (size_t i = 0; i < nreq; ++i)
      if (ps[i].status == JXLB_OK && ps[i].composed) ps[i].layer_info.push_back(ps[i].seen.back());
    // a failed owner drops its layers
    for (size_t t = 0; t < todo.size(); ++t)
      if (ps[todo[t].owner].status != JXLB_OK && ps[first_hidden + t].status == JXLB_OK) Fail(&ps[first_hidden + t], JXLB_ERROR, "owner failed");
  }

  // Host side: parse every request, lay out the batch.  No CUDA calls.
  void Parse(const jxlb_request* reqs, size_t count, int api, const int32_t* frame_index = nullptr) {
    n = count;
    api_level = api <= 0 ? 34 : api;
    ps.assign(n, Parsed());
    // One image's failure stays its own: an allocation failure while parsing (std::bad_alloc is what the reference maps to
    // its "not enough memory" exception, JniDecoding.cpp:81-93) must neither unwind out of a worker thread
    // (std::terminate) nor take the other images of the batch with it.
    ParallelFor(n, [&](size_t i) {
      try {
        ParseRequest(reqs[i], api_level, &ps[i], frame_index ? frame_index[i] : -1);
      } catch (const std::bad_alloc&) {
        ps[i] = Parsed();
        Fail(&ps[i], JXLB_OOM, "Not enough memory to decode this image");
      } catch (const std::exception& e) {
        ps[i] = Parsed();
        Fail(&ps[i], JXLB_ERROR, std::string("Error while decoding: ") + e.what());
      }
    });
    ExpandCompositions(reqs, frame_index);
    const size_t np = ps.size();  // requests + hidden layer entries
    frame_of.assign(np, 0);
    final_off.assign(np, 0);
    final_bytes.assign(np, 0);
    for (size_t i = 0; i < np; ++i) {
      Parsed& p = ps[i];
      if (p.status != JXLB_OK) continue;
      // an image whose planes alone could not be allocated fails by itself instead of failing the batch's allocation
      if (p.plan.work_bytes + p.plan.const_bytes + 3 * p.plan.xyb_bytes + p.plan.up_bytes > kMaxImageDeviceBytes) {
        Fail(&p, JXLB_OOM, "Not enough memory to decode this image");
        continue;
      }
      p.const_off = const_total;
      const_total += Align256(p.plan.const_bytes);
      p.work_off = work_total;
      work_total += Align256(p.plan.work_bytes);
      p.stage_stride = Align256((size_t) p.fw * 4 * (p.out16 ? 2 : 1));
      p.stage_off = stage_total;
      // the fused VarDCT kernel packs straight into final_out; only modular frames (and the unfused debug path) stage RGBA
      if (p.plan.proto.encoding != 0 || UseUnfusedFilters() || p.plan.up_bytes || p.resize || p.color_matrix || p.orient != 1 || p.placed || p.layer_only || p.composed)
        stage_total += Align256(p.stage_stride * p.fh_);
      if (p.composed) {
        p.canvas_bytes = Align256((size_t) p.md.xsize * p.md.ysize * 16);  // float RGBA reference slots; the picture fits in one too
        p.canvas_off = work_total;
        work_total += 5 * p.canvas_bytes;
      }
      if (p.placed) {
        p.place_off = work_total;
        work_total += Align256((size_t) p.md.xsize * p.md.ysize * 4 * (p.out16 ? 2 : 1));
      }
      if (p.orient != 1) {
        p.or_off = work_total;
        work_total += Align256((size_t) p.ow * p.oh * 4 * (p.out16 ? 2 : 1));
      }
      if (p.color_matrix) {
        p.cm_off = const_total;
        const_total += Align256(sizeof(ColorMatrixPlan));
        if (p.cmt16) const_total += Align256(sizeof(ColorMatrixTables16));
        if (p.cmp.tonemap) {  // per-row scratch of the tone mapper (FirstBlackKernel)
          p.cm_rows_off = work_total;
          work_total += Align256((size_t) std::max(p.oh, p.out_h) * 4);
        }
      }
      if (p.resize) {
        auto axis_bytes = [](const ResizeAxis& a) { return Align256(a.start.size() * 4) + Align256(a.count.size() * 4) + Align256(a.weights.size() * 2); };
        p.rs_table_off = const_total;
        p.rs_table_bytes = axis_bytes(p.rp.v) + axis_bytes(p.rp.h);
        const_total += p.rs_table_bytes;
        p.rs_mid_bytes = (p.rp.identity_v || p.rp.nearest) ? 0 : Align256((size_t) p.rp.scaled_h * p.rp.src_w * 4);
        p.rs_scaled_bytes = (p.rp.identity_h && !p.rp.nearest) ? 0 : Align256((size_t) p.rp.scaled_h * p.rp.scaled_w * 4);
        p.rs_work_off = work_total;
        work_total += p.rs_mid_bytes + p.rs_scaled_bytes;
      }
      // a ring slot: one plane set for the fused kernels; two for the unfused filters; + the upsampled planes of a half-resolution frame
      xyb_slot_bytes = std::max(xyb_slot_bytes, Align256(p.plan.xyb_bytes * ((UseUnfusedFilters() || p.plan.up_bytes) ? 2 : 1) + p.plan.up_bytes));
      if (p.plan.proto.encoding == 0) ++n_vardct;
      final_bytes[i] = p.layer_only ? 0 : (size_t) p.out_w * FormatBytesPerPixel((uint32_t) p.format) * p.out_h;
      final_off[i] = final_total;
      final_total += Align256(final_bytes[i]);
      frame_of[i] = nframes++;
      p.status_base = status_total;
      status_total += p.plan.num_streams;
      const FrameDev& f = p.plan.proto;
      // modular streams coded with the frame's global code when that code uses LZ77 (libjxl's effort-1 lossless encoder):
      // every such stream gets a window as large as the symbols it can hold
      bool lz = false;
      if (f.num_mod_channels && p.g.has_global_tree && p.g.tree_code.size() >= sizeof(CodeHeader) &&
          reinterpret_cast<const CodeHeader*>(p.g.tree_code.data())->lz77) {
        const uint64_t side = std::min<uint64_t>(f.group_dim, std::max(f.width, f.height));
        const uint64_t symbols = std::min<uint64_t>(side * side, (uint64_t) f.width * f.height) * f.num_mod_channels + 65536;
        uint32_t e = 1u << 12;
        while (e < symbols && e < (1u << kLz77WindowLog)) e <<= 1;
        // windows are sized for the batch's largest stream: past 8 GiB in total the image goes without (its LZ77 streams
        // then report "unsupported") instead of failing the whole batch's allocation
        const uint64_t img_slots = f.single_section ? 1 : f.num_groups;
        if ((lz_slots + img_slots) * (uint64_t) std::max(lz_entries, e) * 4 <= ((uint64_t) 8 << 30)) {
          lz = true;
          lz_entries = std::max(lz_entries, e);
        }
      }
      if (f.single_section) {
        jobs_single.push_back(StreamJob{frame_of[i], 0, f.num_lf_groups + f.num_groups, lz ? ++lz_slots : 0});
      } else {
        if (f.encoding == 0)
          for (uint32_t l = 0; l < f.num_lf_groups; ++l) jobs_lf.push_back(StreamJob{frame_of[i], l, l, 0});
        // lane-parallel AC decode unless the AC code needs an LZ77 window or does not fit shared memory
        bool lane = false;
        if (f.encoding == 0 && !p.g.ac_code.empty()) {
          const CodeHeader* chh = reinterpret_cast<const CodeHeader*>(p.g.ac_code.data());
          lane = !chh->lz77 && chh->total_bytes <= kMaxAcSmemCode && f.num_passes == 1 && getenv("JXLB_NO_LANE_AC") == nullptr;
          if (lane) {
            ac_smem_code_bytes = std::max(ac_smem_code_bytes, chh->total_bytes);
            if (chh->use_prefix) ac_fast = false;
          }
        }
        if (lane) {
          for (uint32_t g = 0; g < f.num_groups; ++g) jobs_lane_groups.push_back(StreamJob{frame_of[i], g, f.num_lf_groups + g, 0});
          for (uint32_t g0 = 0; g0 < f.num_groups; g0 += AcGroupsPerCta())
            jobs_ac_cta.push_back(AcCtaJob{frame_of[i], g0, std::min(AcGroupsPerCta(), f.num_groups - g0), 0});
          if (f.num_mod_channels > f.global_mod_decoded)
            for (uint32_t g = 0; g < f.num_groups; ++g) jobs_lane_mod.push_back(StreamJob{frame_of[i], g, f.num_lf_groups + g, lz ? ++lz_slots : 0});
        } else {
          for (uint32_t g = 0; g < f.num_groups; ++g) jobs_groups.push_back(StreamJob{frame_of[i], g, f.num_lf_groups + g, lz ? ++lz_slots : 0});
        }
      }
    }
    auto job_bytes = [](const ScratchLayout& l) {
      return Align256(l.arena_bytes) + Align256((size_t) l.wp_ints * 4) + 3 * 1024 + Align256(l.hf_arena_bytes) +
             (l.hf_arena_bytes ? (size_t) 2 * 65536 * 4 : 0);
    };
    sl_single.arena_bytes = 1536u << 10;
    sl_single.wp_ints = ModFastScratch::Ints(1024 + 8);
    sl_single.hf_arena_bytes = 2048u << 10;
    sl_single.max_local_nodes = 32768;
    sl_single.bytes_per_job = Align256(job_bytes(sl_single));
    sl_lf.arena_bytes = 1536u << 10;
    sl_lf.wp_ints = ModFastScratch::Ints(kLfGroupCells * kLfGroupCells + 8);  // BlockInfo channel: up to one entry per cell
    sl_lf.max_local_nodes = 32768;
    sl_lf.bytes_per_job = Align256(job_bytes(sl_lf));
    bool grp_modular = false, grp_local_tree = false;
    uint32_t grp_dim = 256;
    for (size_t i = 0; i < ps.size(); ++i) {
      if (ps[i].status != JXLB_OK || ps[i].plan.proto.single_section) continue;
      const FrameDev& f = ps[i].plan.proto;
      if (f.num_mod_channels > f.global_mod_decoded) {
        grp_modular = true;
        grp_dim = std::max(grp_dim, f.group_dim);
        if (!ps[i].g.has_global_tree) grp_local_tree = true;
      }
    }
    sl_grp.arena_bytes = grp_modular ? (grp_local_tree ? (192u << 10) : (64u << 10)) : 0;  // local tree + code, group-local palette colours
    sl_grp.wp_ints = grp_modular ? ModFastScratch::Ints(grp_dim + 8) : 0;
    sl_grp.max_local_nodes = 2048;
    sl_grp.bytes_per_job = Align256(job_bytes(sl_grp));
    const size_t njobs = jobs_single.size() + jobs_lf.size() + jobs_groups.size() + jobs_lane_groups.size() + jobs_lane_mod.size() +
                         jobs_ac_cta.size();  // AcCtaJob has the same size as StreamJob
    meta_total = Align256(nframes * sizeof(FrameDev)) + Align256(njobs * sizeof(StreamJob));
  }

  bool AnyOk() const {
    for (auto& p : ps)
      if (p.status == JXLB_OK) return true;
    return false;
  }

  // Allocates (or reuses) device buffers, stages the const regions and uploads them.
  void Upload(BatchBuffers* use) {
    buf = use ? use : &own;
    CUDA_OK(cudaSetDevice(ctx->device));
    if (!events) {
      for (auto& set : ev_ring)
        for (auto& e : set) CUDA_OK(cudaEventCreate(&e));
      CUDA_OK(cudaEventCreate(&span_start));
      CUDA_OK(cudaEventCreate(&span_end));
      events = true;
    }
    if (!stream) {
      CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
      CUDA_OK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
      CUDA_OK(cudaStreamCreateWithPriority(&lf_stream, cudaStreamNonBlocking, ctx->lf_priority));
      own_streams = true;
    }
    if (!ev_uploaded) {
      CUDA_OK(cudaEventCreateWithFlags(&ev_uploaded, cudaEventDisableTiming));
      CUDA_OK(cudaEventCreateWithFlags(&ev_lf_done, cudaEventDisableTiming));
    }
    cudaStream_t s = stream;
    const size_t lf_scratch = Align256(sl_lf.bytes_per_job * jobs_lf.size());
    const size_t grp_scratch = Align256(sl_grp.bytes_per_job * std::max(jobs_groups.size(), jobs_lane_mod.size()));
    const size_t single_scratch = Align256(sl_single.bytes_per_job * jobs_single.size());
    buf->const_buf.Ensure(const_total);
    buf->work_buf.Ensure(work_total);
    const size_t lz_scratch = Align256((size_t) lz_slots * lz_entries * 4);
    buf->scratch_buf.Ensure(lf_scratch + grp_scratch + single_scratch + lz_scratch);
    buf->meta_buf.Ensure(meta_total);
    buf->stage_out.Ensure(stage_total);
    buf->final_out.Ensure(final_total);
    buf->xyb_ring.Ensure(xyb_slot_bytes * std::min(kXybRing, std::max<size_t>(n, 1)));
    buf->staging.Ensure(const_total + meta_total);
    buf->status_host.Ensure((size_t) status_total * 4 + 256);
    sl_lf.base = buf->scratch_buf.p;
    sl_grp.base = buf->scratch_buf.p + lf_scratch;
    sl_single.base = buf->scratch_buf.p + lf_scratch + grp_scratch;
    for (ScratchLayout* l : {&sl_lf, &sl_grp, &sl_single}) {
      l->lz_base = lz_slots ? buf->scratch_buf.p + lf_scratch + grp_scratch + single_scratch : nullptr;
      l->lz_entries = lz_entries;
    }
    CUDA_OK(cudaEventRecord(ev_ring[0][0], s));
    uint8_t* stg = buf->staging.p;
    frames.assign(nframes, FrameDev());
    ParallelFor(ps.size(), [&](size_t i) {
      Parsed& p = ps[i];
      if (p.status != JXLB_OK) return;
      FillConstRegion(p.plan, p.cs.data(), p.fh, p.g, stg + p.const_off);
      if (p.color_matrix) {
        memcpy(stg + p.cm_off, &p.cmp, sizeof(ColorMatrixPlan));
        if (p.cmt16) memcpy(stg + p.cm_off + Align256(sizeof(ColorMatrixPlan)), p.cmt16.get(), sizeof(ColorMatrixTables16));
      }
      if (p.resize) {
        uint8_t* t = stg + p.rs_table_off;
        for (const ResizeAxis* a : {&p.rp.v, &p.rp.h}) {
          memcpy(t, a->start.data(), a->start.size() * 4);
          t += Align256(a->start.size() * 4);
          memcpy(t, a->count.data(), a->count.size() * 4);
          t += Align256(a->count.size() * 4);
          memcpy(t, a->weights.data(), a->weights.size() * 2);
          t += Align256(a->weights.size() * 2);
        }
      }
      FrameDev fd = BindFrameDev(p.plan, buf->const_buf.p + p.const_off, buf->work_buf.p + p.work_off);
      if (p.plan.xyb_bytes) {
        uint8_t* slot = buf->xyb_ring.p + (i % kXybRing) * xyb_slot_bytes;
        fd.xyb0 = reinterpret_cast<float*>(slot);
        fd.xyb1 = (UseUnfusedFilters() || p.plan.up_bytes) ? reinterpret_cast<float*>(slot + p.plan.xyb_bytes) : fd.xyb0;
      }
      frames[frame_of[i]] = fd;
    });
    uint8_t* meta_h = stg + const_total;
    memcpy(meta_h, frames.data(), nframes * sizeof(FrameDev));
    StreamJob* jobs_h = reinterpret_cast<StreamJob*>(meta_h + Align256(nframes * sizeof(FrameDev)));
    size_t jo = 0;
    auto put = [&](const std::vector<StreamJob>& v) {
      size_t o = jo;
      if (!v.empty()) memcpy(jobs_h + jo, v.data(), v.size() * sizeof(StreamJob));
      jo += v.size();
      return o;
    };
    const size_t so = put(jobs_single), lo = put(jobs_lf), go = put(jobs_groups), lgo = put(jobs_lane_groups), lmo = put(jobs_lane_mod);
    const size_t aco = jo;
    if (!jobs_ac_cta.empty()) memcpy(jobs_h + jo, jobs_ac_cta.data(), jobs_ac_cta.size() * sizeof(AcCtaJob));
    jo += jobs_ac_cta.size();
    CUDA_OK(cudaMemcpyAsync(buf->const_buf.p, stg, const_total, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(buf->meta_buf.p, meta_h, meta_total, cudaMemcpyHostToDevice, s));
    frames_d = reinterpret_cast<const FrameDev*>(buf->meta_buf.p);
    const StreamJob* jd = reinterpret_cast<const StreamJob*>(buf->meta_buf.p + Align256(nframes * sizeof(FrameDev)));
    jobs_single_d = jd + so;
    jobs_lf_d = jd + lo;
    jobs_groups_d = jd + go;
    jobs_lane_groups_d = jd + lgo;
    jobs_lane_mod_d = jd + lmo;
    jobs_ac_cta_d = reinterpret_cast<const AcCtaJob*>(jd + aco);
    CUDA_OK(cudaEventRecord(ev_ring[0][1], s));
    uploaded = true;
  }

  // Pinned host destinations for the overlapped download (e2e path).  Images that cannot get one fall back to the
  // copy in Finish (which reports JXLB_OOM when that fails too).
  void PrepareHostOutputs() {
    host_dst.assign(n, nullptr);
    img_ev.assign(n, nullptr);
    for (size_t i = 0; i < n; ++i) {
      if (ps[i].status != JXLB_OK) continue;
      host_dst[i] = Pool().Get(final_bytes[i]);
      CUDA_OK(cudaEventCreateWithFlags(&img_ev[i], cudaEventDisableTiming));
    }
  }

  // All kernels, from the uploaded codestreams to packed pixels in HBM.  Asynchronous on this batch's stream.
  void Run() {
    CUDA_OK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream;
    std::lock_guard<std::mutex> enqueue_lock(ctx->mu);  // runs are enqueued one at a time (stage gates, DeviceContext)
    const bool gated = !host_dst.empty() || getenv("JXLB_GATE_PREPARED") != nullptr;
    if (pending_runs >= kEventSets) CollectRuns();  // the ring is full: drain (synchronises)
    if (pending_runs > 0) {                          // keep the upload events of set 0 readable from every set
      run_index = (run_index + 1) % kEventSets;
    }
    ev = ev_ring[run_index];
    ++pending_runs;
    if (span_armed) {
      CUDA_OK(cudaEventRecord(span_start, s));
      span_armed = false;
    }
    CUDA_OK(cudaEventRecord(ev[2], s));
    // everything enqueued on `s` so far (the upload, the previous run) precedes the LF stage
    CUDA_OK(cudaEventRecord(ev_uploaded, s));
    CUDA_OK(cudaStreamWaitEvent(lf_stream, ev_uploaded, 0));
    for (size_t i = 0; i < ps.size(); ++i) {
      Parsed& p = ps[i];
      if (p.status != JXLB_OK) continue;
      uint8_t* wb = buf->work_buf.p + p.work_off;
      CUDA_OK(cudaMemsetAsync(wb + p.plan.off_status, 0xFF, (size_t) p.plan.num_streams * 4, lf_stream));
    }
    {
      const uint64_t k = ctx->gate_lf.Enter(lf_stream, gated);
      LaunchLfGroups(frames_d, jobs_lf_d, (uint32_t) jobs_lf.size(), sl_lf, lf_stream);
      ctx->gate_lf.Leave(k, lf_stream);
    }
    CUDA_OK(cudaEventRecord(ev_lf_done, lf_stream));
    // meanwhile, on the dense stream: clear the coefficient planes
    for (size_t i = 0; i < ps.size(); ++i) {
      Parsed& p = ps[i];
      if (p.status != JXLB_OK) continue;
      uint8_t* wb = buf->work_buf.p + p.work_off;
      if (p.plan.coef_bytes) {
        if (p.plan.coef_bytes % 16 == 0) LaunchFill(wb + p.plan.off_coef, p.plan.coef_bytes, 0u, s);
        else CUDA_OK(cudaMemsetAsync(wb + p.plan.off_coef, 0, p.plan.coef_bytes, s));
      }
    }
    CUDA_OK(cudaStreamWaitEvent(s, ev_lf_done, 0));
    LaunchSingleSectionFrames(frames_d, jobs_single_d, (uint32_t) jobs_single.size(), ctx->nat_dev, sl_single, s);
    CUDA_OK(cudaEventRecord(ev[3], s));
    const uint64_t k_ac = ctx->gate_ac.Enter(s, gated);
    LaunchPassGroups(frames_d, jobs_groups_d, (uint32_t) jobs_groups.size(), ctx->nat_dev, sl_grp, s);
    LaunchBuildGroupBlocks(frames_d, jobs_lane_groups_d, (uint32_t) jobs_lane_groups.size(), s);
    LaunchAcLanes(frames_d, jobs_ac_cta_d, (uint32_t) jobs_ac_cta.size(), ctx->nat_dev, ac_smem_code_bytes, ac_fast, s);
    LaunchGroupModular(frames_d, jobs_lane_mod_d, (uint32_t) jobs_lane_mod.size(), sl_grp, s);
    LaunchFrameStatus(frames_d, nframes, s);
    ctx->gate_ac.Leave(k_ac, s);
    CUDA_OK(cudaEventRecord(ev[4], s));
    const uint64_t k_recon = ctx->gate_recon.Enter(s, gated);
    EnsureSampleEvents();
    cudaEvent_t* sev = sample_ev[run_index].data();
    size_t sampled = 0, vd = 0;
    // hidden layer entries first: their pictures must exist when the owners composite them (one stream: enqueue order)
    for (size_t oi = 0; oi < ps.size(); ++oi) {
      const size_t i = oi < ps.size() - n ? n + oi : oi - (ps.size() - n);
      Parsed& p = ps[i];
      if (p.status != JXLB_OK) continue;
      const FrameDev& f = frames[frame_of[i]];
      OutputDesc od;
      od.data = buf->stage_out.p + p.stage_off;
      od.stride_bytes = (uint32_t) p.stage_stride;
      od.bits16 = p.out16;
      od.alpha_channel = -1;
      od.alpha_bits = 8;
      od.color_bits = p.md.bits_per_sample;
      int ai = p.md.alpha_channel();
      if (ai >= 0) {
        od.alpha_channel = (int32_t) (f.num_color_mod_channels + (uint32_t) ai);
        od.alpha_bits = p.md.extra[ai].bits;
      }
      PackParams pk;  // ReformatColorConfig
      pk.src = od.data;
      pk.src_stride = od.stride_bytes;
      pk.width = p.fw;
      pk.height = p.fh_;
      pk.src16 = p.out16;
      pk.depth = p.depth;
      pk.format = (uint32_t) p.format;
      pk.associate = (!p.alpha_premultiplied && p.has_alpha) ? 1 : 0;
      pk.attenuate = !p.alpha_premultiplied ? 1 : 0;
      pk.dst_stride = pk.width * FormatBytesPerPixel(pk.format);
      pk.dst = buf->final_out.p + final_off[i];
      const PackParams pk_final = pk;
      const bool post = p.resize || p.color_matrix || p.orient != 1 || p.placed || p.layer_only || p.composed;
      if (post) {  // the decode stage hands straight RGBA8 to the rescaler / colour pass; ReformatColorConfig runs on their result
        pk.format = 4;  // staging: samples as decoded (RGBA8, or RGBA16 for sources deeper than 8 bits)
        pk.associate = 0;
        pk.dst = od.data;
        pk.dst_stride = od.stride_bytes;
      }
      if (f.sq_nch) LaunchUnsqueeze(f, p.g.sq.steps.data(), s);
      if (f.global_planes) LaunchScatterGlobalPlanes(f, s);
      if (f.encoding == 0 && !f.single_section && !f.sq_nch && f.num_mod_channels && f.global_nb_transforms) LaunchModularGlobalInverse(f, s);
      if (f.encoding == 0) {
        const bool timed = (vd++ % kTimeEvery) == 0;
        LaunchLfFinal(f, s);
        if (timed) CUDA_OK(cudaEventRecord(sev[3 * sampled], s));
        LaunchRecon(f, ctx->nt_dev, s);
        if (timed) CUDA_OK(cudaEventRecord(sev[3 * sampled + 1], s));
        if (f.upsampling == 2) {
          // half-resolution frame: separate filter kernels, 2x upsampling of the XYB planes, colour + pack at full size
          const int cur = LaunchFilters(f, ctx->nt_dev, s);
          float* up = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(f.xyb0) + 2 * p.plan.xyb_bytes);
          LaunchUpsample2(f, cur ? f.xyb1 : f.xyb0, up, f.up_stride, f.up_h, s);
          FrameDev fu = f;
          fu.width = f.up_width;
          fu.height = f.up_height;
          fu.plane_stride = f.up_stride;
          fu.plane_h = f.up_h;
          OutputDesc odu = od;
          if (od.alpha_channel >= 0) {  // the alpha plane is coded at half resolution too
            int32_t* up_alpha = reinterpret_cast<int32_t*>(up + (size_t) 3 * f.up_h * f.up_stride);
            LaunchUpsampleAlpha2(f, f.mod + (size_t) od.alpha_channel * f.height * f.mod_stride, od.alpha_bits, up_alpha, f.up_stride, s);
            fu.mod = up_alpha;
            fu.mod_stride = f.up_stride;
            odu.alpha_channel = 0;
            odu.alpha_float = 1;
          }
          LaunchColor(fu, p.cp, ctx->nt_dev, up, odu, s);
          LaunchPack(pk, s);
        } else if (!UseUnfusedFilters()) {
          LaunchFilterColorPack(f, p.cp, ctx->nt_dev, od, pk, s);
        } else {
          int cur = LaunchFilters(f, ctx->nt_dev, s);
          LaunchColor(f, p.cp, ctx->nt_dev, cur ? f.xyb1 : f.xyb0, od, s);
          LaunchPack(pk, s);
        }
        if (timed) {
          CUDA_OK(cudaEventRecord(sev[3 * sampled + 2], s));
          ++sampled;
        }
      } else {
        if (!f.single_section && f.global_serial) LaunchModularGlobalInverse(f, s);  // delta palette: serial inverse first
        LaunchModularToRgba(f, od, s);
        if (!post) LaunchPack(pk, s);
      }
      if (p.layer_only) continue;  // a layer: its straight RGBA picture stays in the staging buffer for its owner
      const uint8_t* res = od.data;
      uint32_t res_stride = od.stride_bytes;
      if (p.composed) {
        // reference slots 0 .. 3 and the current picture; every layer is blended over the slot it names and stored in
        // the slot it is saved to (blending reads and writes a pixel in the same thread, so source == destination is fine)
        const uint32_t bpp = p.out16 ? 8 : 4;
        uint8_t* canvas = buf->work_buf.p + p.canvas_off;
        float4* slot[4] = {nullptr, nullptr, nullptr, nullptr};
        uint8_t* cur = canvas + 4 * p.canvas_bytes;
        for (size_t l = 0; l < p.layer_info.size(); ++l) {
          const Parsed::SeenFrame& sf = p.layer_info[l];
          const bool own = l + 1 == p.layer_info.size();
          const Parsed& lp = own ? p : ps[p.layers[l]];
          CompositeParams cpz{};
          cpz.bg = slot[sf.blend.source & 3];
          cpz.fg = buf->stage_out.p + lp.stage_off;
          cpz.fg_stride = (uint32_t) lp.stage_stride;
          cpz.fw = lp.fw;
          cpz.fh = lp.fh_;
          cpz.x0 = sf.x0;
          cpz.y0 = sf.y0;
          cpz.out = own ? cur : nullptr;
          cpz.save = (!own && sf.saves) ? reinterpret_cast<float4*>(canvas + sf.slot * p.canvas_bytes) : nullptr;
          cpz.cw = p.md.xsize;
          cpz.ch = p.md.ysize;
          cpz.canvas_stride = p.md.xsize * bpp;
          cpz.bits16 = p.out16;
          cpz.has_alpha = p.has_alpha;
          cpz.alpha_premultiplied = p.alpha_premultiplied;
          cpz.mode_color = sf.blend.mode;
          cpz.mode_alpha = sf.alpha_blend.mode;
          cpz.clamp = sf.blend.clamp || sf.alpha_blend.clamp;
          cpz.dither = (own && !p.out16) ? reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(ctx->nt_dev) + offsetof(NumericTables, dither)) : nullptr;
          cpz.orientation = p.md.orientation;
          LaunchComposite(cpz, s);
          if (cpz.save) slot[sf.slot] = cpz.save;
        }
        res = cur;
        res_stride = p.md.xsize * bpp;
      }
      const ColorMatrixPlan* cm_dev = reinterpret_cast<const ColorMatrixPlan*>(buf->const_buf.p + p.cm_off);
      if (p.placed) {
        uint8_t* dst = buf->work_buf.p + p.place_off;
        const uint32_t bpp = p.out16 ? 8 : 4;
        LaunchPlace(od.data, od.stride_bytes, p.fw, p.fh_, bpp, p.place_x0, p.place_y0, p.has_alpha ? 0u : (p.out16 ? 0xFFFFu : 0xFFu), dst,
                    p.md.xsize * bpp, p.md.xsize, p.md.ysize, s);
        res = dst;
        res_stride = p.md.xsize * bpp;
      }
      if (p.orient != 1) {
        uint8_t* dst = buf->work_buf.p + p.or_off;
        const uint32_t bpp = p.out16 ? 8 : 4;
        LaunchOrient(res, res_stride, p.md.xsize, p.md.ysize, bpp, p.orient, dst, p.ow * bpp, s);
        res = dst;
        res_stride = p.ow * bpp;
      }
      uint32_t* cm_rows = reinterpret_cast<uint32_t*>(buf->work_buf.p + p.cm_rows_off);
      if (p.color_matrix && p.matrix_before_resize)
        LaunchColorMatrix(const_cast<uint8_t*>(res), res_stride, p.ow, p.oh, cm_dev, p.cmp.tonemap != 0, p.out16, cm_rows, s);
      if (p.resize) {
        ResizeDev rd{};
        rd.src = res;
        rd.src_stride = res_stride;
        rd.src_w = p.rp.src_w;
        rd.src_h = p.rp.src_h;
        rd.scaled_w = p.rp.scaled_w;
        rd.scaled_h = p.rp.scaled_h;
        rd.has_v = !p.rp.identity_v;
        rd.has_h = !p.rp.identity_h;
        rd.nearest = p.rp.nearest;
        rd.premultiply = p.rp.premultiply;
        const uint8_t* t = buf->const_buf.p + p.rs_table_off;
        for (int ax = 0; ax < 2; ++ax) {
          const ResizeAxis& a = ax ? p.rp.h : p.rp.v;
          ResizeAxisDev& d = ax ? rd.h : rd.v;
          d.start = reinterpret_cast<const uint32_t*>(t);
          t += Align256(a.start.size() * 4);
          d.count = reinterpret_cast<const uint32_t*>(t);
          t += Align256(a.count.size() * 4);
          d.weights = reinterpret_cast<const int16_t*>(t);
          t += Align256(a.weights.size() * 2);
          d.taps = a.taps;
        }
        rd.mid = buf->work_buf.p + p.rs_work_off;
        rd.scaled = rd.mid + p.rs_mid_bytes;
        res = LaunchResize(rd, s);
        res_stride = res == rd.scaled ? rd.scaled_w * 4 : res == rd.mid ? rd.src_w * 4 : rd.src_stride;
        if (p.rp.zero_tail_rows)  // horizontal-only quirk of pic-scale 0.7.6 (resize.h): the last height % 4 rows stay zero
          CUDA_OK(cudaMemsetAsync(rd.scaled + (size_t) (rd.scaled_h - p.rp.zero_tail_rows) * rd.scaled_w * 4, 0,
                                  (size_t) p.rp.zero_tail_rows * rd.scaled_w * 4, s));
        res += (size_t) p.rp.crop_y * res_stride + (size_t) p.rp.crop_x * 4;
      }
      if (post) {
        PackParams pr = pk_final;
        pr.src = res;
        pr.src_stride = res_stride;
        pr.width = p.out_w;
        pr.height = p.out_h;
        pr.dst_stride = pr.width * FormatBytesPerPixel(pr.format);
        if (p.resize && p.rp.zero_last_row) {  // pic-scale 0.7.6 crop quirk (resize.h): the last row of a column-cropped picture is zero
          --pr.height;
          CUDA_OK(cudaMemsetAsync(pr.dst + (size_t) pr.height * pr.dst_stride, 0, pr.dst_stride, s));
          // ReformatColorConfig of a zero row is a zero row in every target format
        }
        // colour pass after the rescale, before the reformat (JniDecoding.cpp:120-228); the zero rows of the quirks above
        // pass through it unchanged only when the tables map 0 to 0, which they do (toLinear(0) = 0, gamma[0] = 0)
        if (p.color_matrix && !p.matrix_before_resize && pr.height)
          LaunchColorMatrix(const_cast<uint8_t*>(res), res_stride, pr.width, pr.height, cm_dev, p.cmp.tonemap != 0, p.out16, cm_rows, s);
        if (pr.height) LaunchPack(pr, s);
      }
      // download: only the "image done" event is recorded here; Finish() enqueues each copy once its image is complete
      if (i < host_dst.size() && host_dst[i]) CUDA_OK(cudaEventRecord(img_ev[i], s));
    }
    sampled_in_run[run_index] = sampled;
    ctx->gate_recon.Leave(k_recon, s);
    CUDA_OK(cudaEventRecord(ev[7], s));
    CUDA_OK(cudaEventRecord(span_end, s));
    ran = true;
  }

  // Waits for every run issued so far and reads its stage times (stage_ms = the last run, stage_sum / stage_runs = all).
  // ms: [0] upload, [1] LF sections, [2] group sections, [3] LF final, [4] inverse transforms, [5] filters+colour+pack,
  //     [6] download, [7] all kernels
  void CollectRuns() {
    CUDA_OK(WaitStream(stream));
    for (int k = pending_runs - 1; k >= 0; --k) {
      const int set = ((run_index - k) % kEventSets + kEventSets) % kEventSets;
      cudaEvent_t* e = ev_ring[set];
      auto el = [&](cudaEvent_t a, cudaEvent_t b) {
        float v = 0;
        if (cudaEventElapsedTime(&v, a, b) != cudaSuccess) {
          cudaGetLastError();
          v = 0;
        }
        return v;
      };
      stage_ms[0] = upload_timed ? 0.f : el(ev_ring[0][0], ev_ring[0][1]);
      upload_timed = true;
      stage_ms[1] = el(e[2], e[3]);
      stage_ms[2] = el(e[3], e[4]);
      // reconstruction runs image by image: [3] = the whole phase, [4] / [5] = mean sampled kernel time x images
      stage_ms[3] = el(e[4], e[7]);
      {
        double idct = 0, filt = 0;
        const size_t ns = sampled_in_run[set];
        for (size_t k = 0; k < ns; ++k) {
          idct += el(sample_ev[set][3 * k], sample_ev[set][3 * k + 1]);
          filt += el(sample_ev[set][3 * k + 1], sample_ev[set][3 * k + 2]);
        }
        stage_ms[4] = ns ? (float) (idct / ns * n_vardct) : 0.f;
        stage_ms[5] = ns ? (float) (filt / ns * n_vardct) : 0.f;
      }
      stage_ms[6] = finished[set] ? el(e[7], e[8]) : 0.f;
      stage_ms[7] = el(e[2], e[7]);
      finished[set] = false;
      if (ctx->timeline)
        fprintf(stderr, "[timeline] batch %p run: start %.2f lf_done %.2f groups_done %.2f recon+filter_done %.2f ms\n", (void*) this,
                el(ctx->origin, e[2]), el(ctx->origin, e[3]), el(ctx->origin, e[4]), el(ctx->origin, e[7]));
      for (int i = 0; i < 8; ++i) stage_sum[i] += stage_ms[i];
      ++stage_runs;
    }
    pending_runs = 0;
  }

  cudaStream_t D2hStream() const {
    static const bool shared = getenv("JXLB_D2H_SHARED") != nullptr;
    return shared ? ctx->d2h_stream : copy_stream;
  }
  // Downloads the per-stream statuses (and, if host_out, the pixels), synchronises and resolves per-image status.
  void Finish(bool to_host, int output_device, std::vector<DecodedImage>* out) {
    cudaStream_t s = stream;
    // Overlapped download (e2e path).  The calling thread would only block in a synchronize anyway, so it waits for each
    // image's "done" event and enqueues that image's copy THEN.  Copies queued ahead of time with cudaStreamWaitEvent
    // sit at the head of the copy engine's queue until their kernels finish and block every copy behind them --
    // including the codestream uploads of the other batches in flight, whose kernels then start ~200 ms late
    // (profiles/r1c_overlap_notes.txt).
    if (to_host)
      for (size_t i = 0; i < n && i < host_dst.size(); ++i) {
        if (!host_dst[i] || ps[i].status != JXLB_OK) continue;
        CUDA_OK(WaitEvent(img_ev[i]));
        CUDA_OK(cudaMemcpyAsync(host_dst[i], buf->final_out.p + final_off[i], final_bytes[i], cudaMemcpyDeviceToHost, D2hStream()));
      }
    if (ran) CUDA_OK(WaitEvent(ev[7]));  // every kernel of the run is done: the status words are final
    uint32_t* sh = reinterpret_cast<uint32_t*>(buf->status_host.p);
    for (size_t i = 0; i < ps.size(); ++i) {
      Parsed& p = ps[i];
      if (p.status != JXLB_OK) continue;
      CUDA_OK(cudaMemcpyAsync(sh + p.status_base, buf->work_buf.p + p.work_off + p.plan.off_status, (size_t) p.plan.num_streams * 4,
                              cudaMemcpyDeviceToHost, s));
    }
    std::vector<void*> result(n, nullptr);
    for (size_t i = 0; i < n && out; ++i) {
      if (ps[i].status != JXLB_OK) continue;
      if (to_host && i < host_dst.size() && host_dst[i]) {
        result[i] = host_dst[i];  // in flight on the copy stream (above)
        host_dst[i] = nullptr;
      } else if (to_host) {
        result[i] = Pool().Get(final_bytes[i]);
        if (!result[i]) {
          Fail(&ps[i], JXLB_OOM, "Not enough memory to decode this image");
          continue;
        }
        CUDA_OK(cudaMemcpyAsync(result[i], buf->final_out.p + final_off[i], final_bytes[i], cudaMemcpyDeviceToHost, s));
      } else {
        void* d = nullptr;
        if (cudaMalloc(&d, final_bytes[i]) != cudaSuccess) {
          Fail(&ps[i], JXLB_OOM, "Not enough memory to decode this image");
          cudaGetLastError();
          continue;
        }
        result[i] = d;
        CUDA_OK(cudaMemcpyAsync(d, buf->final_out.p + final_off[i], final_bytes[i], cudaMemcpyDeviceToDevice, s));
      }
    }
    if (to_host) {  // the decode stream's timeline ends when the copy stream has drained
      CUDA_OK(cudaEventRecord(ev[9], D2hStream()));
      CUDA_OK(cudaStreamWaitEvent(s, ev[9], 0));
    }
    CUDA_OK(cudaEventRecord(ev[8], s));
    finished[run_index] = true;
    CollectRuns();
    tm.ms[0] = stage_ms[0];
    tm.ms[1] = stage_ms[1] + stage_ms[2];
    tm.ms[2] = stage_ms[4];
    tm.ms[3] = stage_ms[5];
    tm.ms[4] = stage_ms[6];
    tm.ms[5] = stage_ms[7] + stage_ms[6];
    const int32_t* shs = reinterpret_cast<const int32_t*>(buf->status_host.p);
    // hidden layer entries first (indices n ..), so that an owner sees its layers' failures
    for (size_t oi = 0; oi < ps.size(); ++oi) {
      const size_t i = oi < ps.size() - n ? n + oi : oi - (ps.size() - n);
      Parsed& p = ps[i];
      auto drop = [&]() {
        if (i >= n || !result[i]) return;
        if (to_host) Pool().Put(result[i]);
        else cudaFree(result[i]);
        result[i] = nullptr;
      };
      if (p.status == JXLB_OK)
        for (size_t l : p.layers)
          if (ps[l].status != JXLB_OK) {
            Fail(&p, ps[l].status, ps[l].message);
            break;
          }
      if (p.status != JXLB_OK) {
        drop();
        continue;
      }
      const FrameDev& f = p.plan.proto;
      int worst = kOk;
      auto check = [&](uint32_t slot) {
        int v = shs[p.status_base + slot];
        if (v == -1) v = kErrBadStream;  // never written
        if (v != kOk && worst == kOk) worst = v;
      };
      if (f.single_section) {
        check(f.num_lf_groups + f.num_groups);
      } else {
        if (f.encoding == 0)
          for (uint32_t l = 0; l < f.num_lf_groups; ++l) check(l);
        for (uint32_t g = 0; g < f.num_groups; ++g) check(f.num_lf_groups + g);
      }
      if (worst != kOk) {
        if (worst == kErrUnsupported || worst == kErrScratch)
          Fail(&p, JXLB_UNSUPPORTED, "coding tool or stream size outside this build's coverage");
        else
          Fail(&p, JXLB_INVALID_JXL, worst == kErrTruncated ? "truncated section" : "corrupt section");
        drop();
        continue;
      }
      if (!out || i >= n) continue;
      DecodedImage& d = (*out)[i];
      d.width = p.out_w;
      d.height = p.out_h;
      d.stride_bytes = p.out_w * FormatBytesPerPixel((uint32_t) p.format);
      d.format = p.format;
      d.color_space = p.color_space;
      d.premultiplied = p.has_alpha ? 1 : 0;
      d.data = result[i];
      d.device = to_host ? -1 : ctx->device;
      (void) output_device;
    }
  }

  float stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

void FreeImageMemory(void* data, int device) {
  if (!data) return;
  if (device < 0) {
    if (!Pool().Put(data)) cudaFreeHost(data);
  } else {
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(device);
    cudaFree(data);
    cudaSetDevice(cur);
  }
}

namespace {
// Picks a decode slot: the lowest-numbered free one (a lone caller therefore always reuses slot 0 and its warm buffers;
// concurrent callers spread over the slots), else waits for one in round-robin order.
Slot* AcquireSlot(DeviceContext* ctx, std::unique_lock<std::mutex>* lock) {
  for (int k = 0; k < kSlots; ++k) {
    Slot* sl = &ctx->slots[k];
    std::unique_lock<std::mutex> l(sl->mu, std::try_to_lock);
    if (l.owns_lock()) {
      *lock = std::move(l);
      return sl;
    }
  }
  Slot* sl = &ctx->slots[ctx->next_slot.fetch_add(1) % kSlots];
  *lock = std::unique_lock<std::mutex>(sl->mu);
  return sl;
}
void CopyStatuses(const Batch& b, std::vector<DecodedImage>* out, int* overall) {
  for (size_t i = 0; i < b.n; ++i) {
    (*out)[i].status = b.ps[i].status;
    (*out)[i].message = b.ps[i].message;
    if (b.ps[i].status != JXLB_OK) *overall = b.ps[i].status;
  }
}
}  // namespace

// Device half of a decode call: slot, upload, kernels, downloads, per-image statuses.  `b` has been parsed.
static int DeviceDecode(Batch& b, int device, int output_device, std::vector<DecodedImage>* out, BatchTimings* timings, double parse_ms) {
  using Clock = std::chrono::steady_clock;
  const auto t0 = Clock::now();
  auto ms_since = [&](Clock::time_point a) { return std::chrono::duration<double, std::milli>(Clock::now() - a).count(); };
  int overall = JXLB_OK;
  if (!b.AnyOk()) {
    CopyStatuses(b, out, &overall);
    return overall;
  }
  try {
    b.ctx = GetContext(device);
  } catch (CudaError& e) {
    for (auto& p : b.ps)
      if (p.status == JXLB_OK) Fail(&p, JXLB_ERROR_NO_DEVICE, e.msg);
    CopyStatuses(b, out, &overall);
    return overall;
  }
  const auto t1 = Clock::now();
  std::unique_lock<std::mutex> lock;
  Slot* slot = AcquireSlot(b.ctx, &lock);
  const double slot_ms = ms_since(t1);
  b.stream = slot->stream;
  b.copy_stream = slot->copy_stream;
  b.lf_stream = slot->lf_stream;
  try {
    const auto t2 = Clock::now();
    b.Upload(&slot->spare);
    if (output_device < 0) b.PrepareHostOutputs();
    const double upload_ms = ms_since(t2);
    const auto t3 = Clock::now();
    b.Run();
    const double launch_ms = ms_since(t3);
    const auto t4 = Clock::now();
    b.Finish(output_device < 0, output_device, out);
    if (b.ctx->timeline)
      fprintf(stderr, "[host] decode_batch n=%zu: parse %.1f  slot wait %.1f  stage+upload %.1f  launch %.1f  finish(wait) %.1f  total %.1f ms | device: "
              "upload %.1f lf %.1f groups %.1f recon-phase %.1f download-tail %.1f kernels %.1f\n", b.n,
              parse_ms, slot_ms, upload_ms, launch_ms, ms_since(t4), parse_ms + ms_since(t0), b.stage_ms[0], b.stage_ms[1], b.stage_ms[2], b.stage_ms[3],
              b.stage_ms[6], b.stage_ms[7]);
    if (timings) *timings = b.tm;
  } catch (CudaError& e) {
    for (auto& p : b.ps)
      if (p.status == JXLB_OK) Fail(&p, JXLB_ERROR_NO_DEVICE, e.msg);
    cudaGetLastError();
  } catch (const std::bad_alloc&) {
    for (auto& p : b.ps)
      if (p.status == JXLB_OK) Fail(&p, JXLB_OOM, "Not enough memory to decode this image");
  } catch (const std::exception& e) {
    for (auto& p : b.ps)
      if (p.status == JXLB_OK) Fail(&p, JXLB_ERROR, std::string("Error while decoding: ") + e.what());
  }
  CopyStatuses(b, out, &overall);
  return overall;
}

int DecodeBatch(const jxlb_request* reqs, size_t n, int api_level, int device, int output_device, std::vector<DecodedImage>* out,
                BatchTimings* timings, const int32_t* frame_index) {
  using Clock = std::chrono::steady_clock;
  const auto t0 = Clock::now();
  out->assign(n, DecodedImage());
  Batch b;
  b.Parse(reqs, n, api_level, frame_index);
  const double parse_ms = std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
  return DeviceDecode(b, device, output_device, out, timings, parse_ms);
}

// ---- submit / collect: the same decode, with the device half on a worker thread -----------------------------------------
// One synchronous call is a latency chain (parse -> upload -> LF stage, ~86 ms of serial entropy chains whatever the batch
// size -> AC -> reconstruction overlapped with the downloads): alone it leaves the GPU and the PCIe link idle most of the
// time.  The reference's entry points are re-entrant and its callers (Glide / Coil worker pools) overlap calls; a caller
// with ONE thread gets the same overlap by submitting a few batches and collecting them in order.  The input buffers
// are only read inside SubmitBatch (parse copies the codestreams, as the JNI entry copies its byte array on entry).
struct PendingBatch {
  Batch b;
  std::vector<DecodedImage> out;
  BatchTimings tm;
  int rc = JXLB_OK;
  std::thread worker;
};

PendingBatch* SubmitBatch(const jxlb_request* reqs, size_t n, int api_level, int device, int output_device, const int32_t* frame_index) {
  using Clock = std::chrono::steady_clock;
  const auto t0 = Clock::now();
  std::unique_ptr<PendingBatch> p(new PendingBatch());
  p->out.assign(n, DecodedImage());
  p->b.Parse(reqs, n, api_level, frame_index);
  const double parse_ms = std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
  PendingBatch* raw = p.get();
  p->worker = std::thread([raw, device, output_device, parse_ms]() { raw->rc = DeviceDecode(raw->b, device, output_device, &raw->out, &raw->tm, parse_ms); });
  return p.release();
}

int CollectBatch(PendingBatch* p, std::vector<DecodedImage>* out, BatchTimings* timings) {
  if (!p) return JXLB_BAD_ARG;
  if (p->worker.joinable()) p->worker.join();
  *out = std::move(p->out);
  if (timings) *timings = p->tm;
  const int rc = p->rc;
  delete p;
  return rc;
}

// ---- prepared batches (throughput interface: inputs resident in HBM, results left in HBM) --------------------------------
Batch* PrepareBatch(const jxlb_request* reqs, size_t n, int api_level, int device, std::vector<int>* status) {
  std::unique_ptr<Batch> b(new Batch());
  b->Parse(reqs, n, api_level);
  status->assign(n, 0);
  try {
    b->ctx = GetContext(device);
    if (b->AnyOk()) {
      b->Upload(nullptr);
      CUDA_OK(cudaStreamSynchronize(b->stream));
    }
  } catch (CudaError& e) {
    for (auto& p : b->ps)
      if (p.status == JXLB_OK) Fail(&p, JXLB_ERROR_NO_DEVICE, e.msg);
    cudaGetLastError();
  }
  for (size_t i = 0; i < n; ++i) (*status)[i] = b->ps[i].status;
  return b.release();
}

int RunBatch(Batch* b, bool sync) {
  if (!b || !b->uploaded) return JXLB_ERROR;
  std::lock_guard<std::mutex> lock(b->mu);
  try {
    b->Run();
    if (sync) b->Finish(false, 0, nullptr);
  } catch (CudaError& e) {
    cudaGetLastError();
    return JXLB_ERROR_NO_DEVICE;
  }
  for (auto& p : b->ps)
    if (p.status != JXLB_OK) return p.status;
  return JXLB_OK;
}

int FetchBatchImage(Batch* b, size_t i, DecodedImage* out) {
  if (!b || i >= b->n || !b->ran) return JXLB_BAD_ARG;
  const Parsed& p = b->ps[i];
  out->status = p.status;
  out->message = p.message;
  if (p.status != JXLB_OK) return p.status;
  std::lock_guard<std::mutex> lock(b->mu);
  void* h = Pool().Get(b->final_bytes[i]);
  if (!h) return JXLB_OOM;
  cudaSetDevice(b->ctx->device);
  if (cudaMemcpyAsync(h, b->buf->final_out.p + b->final_off[i], b->final_bytes[i], cudaMemcpyDeviceToHost, b->stream) != cudaSuccess ||
      cudaStreamSynchronize(b->stream) != cudaSuccess) {
    Pool().Put(h);
    return JXLB_ERROR_NO_DEVICE;
  }
  out->data = h;
  out->device = -1;
  out->width = p.out_w;
  out->height = p.out_h;
  out->stride_bytes = p.out_w * FormatBytesPerPixel((uint32_t) p.format);
  out->format = p.format;
  out->color_space = p.color_space;
  out->premultiplied = p.has_alpha ? 1 : 0;
  return JXLB_OK;
}

void BatchStageMs(const Batch* b, float* ms8) {
  for (int i = 0; i < 8; ++i) ms8[i] = b ? b->stage_ms[i] : 0.f;
}
const void* BatchDevicePixels(const Batch* b, size_t i, size_t* bytes) {
  if (!b || i >= b->n || b->ps[i].status != JXLB_OK) return nullptr;
  if (bytes) *bytes = b->final_bytes[i];
  return b->buf->final_out.p + b->final_off[i];
}
cudaStream_t BatchStream(const Batch* b) { return b->stream; }

// Waits for every asynchronous run of the batch (RunBatch(b, false)) and resolves its status like a synchronous run.
int WaitBatch(Batch* b) {
  if (!b || !b->uploaded || !b->ran) return JXLB_ERROR;
  std::lock_guard<std::mutex> lock(b->mu);
  try {
    b->Finish(false, 0, nullptr);
  } catch (CudaError& e) {
    cudaGetLastError();
    return JXLB_ERROR_NO_DEVICE;
  }
  for (auto& p : b->ps)
    if (p.status != JXLB_OK) return p.status;
  return JXLB_OK;
}
void ResetBatchStats(Batch* b) {
  if (!b) return;
  std::lock_guard<std::mutex> lock(b->mu);
  for (double& v : b->stage_sum) v = 0;
  b->stage_runs = 0;
  b->span_armed = true;
}
// Device time from the start of `first`'s first run since its last ResetBatchStats to the end of `last`'s latest run
// (both must have completed: call WaitBatch on them first).  Negative on error.
float BatchSpanMs(const Batch* first, const Batch* last) {
  if (!first || !last || !first->span_start || !last->span_end) return -1.f;
  float v = 0;
  if (cudaEventElapsedTime(&v, first->span_start, last->span_end) != cudaSuccess) {
    cudaGetLastError();
    return -1.f;
  }
  return v;
}
void BatchStageMsMean(const Batch* b, float* ms8, int* runs) {
  for (int i = 0; i < 8; ++i) ms8[i] = (b && b->stage_runs) ? (float) (b->stage_sum[i] / b->stage_runs) : 0.f;
  if (runs) *runs = b ? b->stage_runs : 0;
}
void FreeBatch(Batch* b) { delete b; }


// ---- test hook: one synthetic block through the reconstruction kernels (test_block.h) ---------------------------------
int TestReconBlock(int device, uint32_t strategy, const int16_t* q, const float* lf, uint32_t hf_mul, uint32_t global_scale,
                   float* out) {
  if (strategy >= (uint32_t) kNumStrategies || !q || !lf || !out || !hf_mul || !global_scale) return JXLB_BAD_ARG;
  uint8_t *cb_d = nullptr, *wb_d = nullptr;
  float* xyb_d = nullptr;
  cudaStream_t st = nullptr;
  int rc = JXLB_OK;
  try {
    DeviceContext* ctx = GetContext(device);
    CUDA_OK(cudaSetDevice(ctx->device));
    ImageMetadata md;
    FrameHeader fh;
    FrameGlobals g;
    FramePlan plan;
    MakeSingleBlockPlan(strategy, global_scale, &md, &fh, &g, &plan);
    std::vector<uint8_t> cb(plan.const_bytes), wb(plan.work_bytes), cs(16);
    FillConstRegion(plan, cs.data(), fh, g, cb.data());
    FillSingleBlock(BindFrameDev(plan, cb.data(), wb.data()), strategy, q, lf, hf_mul);
    CUDA_OK(cudaMalloc(&cb_d, plan.const_bytes));
    CUDA_OK(cudaMalloc(&wb_d, plan.work_bytes));
    CUDA_OK(cudaMalloc(&xyb_d, plan.xyb_bytes));
    CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CUDA_OK(cudaMemcpyAsync(cb_d, cb.data(), plan.const_bytes, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(wb_d, wb.data(), plan.work_bytes, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemsetAsync(xyb_d, 0, plan.xyb_bytes, st));
    FrameDev f = BindFrameDev(plan, cb_d, wb_d);
    f.xyb0 = f.xyb1 = xyb_d;
    LaunchRecon(f, ctx->nt_dev, st);
    const uint32_t R = 8 * StrategyCellsY(strategy), C = 8 * StrategyCellsX(strategy);
    CUDA_OK(cudaMemcpy2DAsync(out, C * sizeof(float), xyb_d, f.plane_stride * sizeof(float), C * sizeof(float), R, cudaMemcpyDeviceToHost, st));
    for (uint32_t c = 1; c < 3; ++c)
      CUDA_OK(cudaMemcpy2DAsync(out + (size_t) c * R * C, C * sizeof(float), xyb_d + (size_t) c * f.plane_h * f.plane_stride,
                                f.plane_stride * sizeof(float), C * sizeof(float), R, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
  } catch (CudaError&) {
    cudaGetLastError();
    rc = JXLB_ERROR_NO_DEVICE;
  } catch (const std::exception&) {
    rc = JXLB_OOM;
  }
  if (st) cudaStreamDestroy(st);
  cudaFree(cb_d);
  cudaFree(wb_d);
  cudaFree(xyb_d);
  return rc;
}

}  // namespace jxlb
