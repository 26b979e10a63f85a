// JxlAnimatedImage entry points (include/jxlb200.h).  The frame table (count, durations, loops, size) is parsed on the
// CPU exactly as the reference's JxlAnimatedDecoder constructor does
// (/root/reference/jxlcoder/src/main/cpp/interop/JxlAnimatedDecoder.hpp:68-185).  getFrame(i) decodes displayed frame i
// on the GPU through the still-image pipeline: the frames the reference's own JxlAnimatedEncoder writes are full-canvas
// kReplace frames (interop/JxlAnimatedEncoder.hpp:111-118), i.e. independent pictures, so frame i needs no earlier
// frame (which is also what lets a 120-frame animation spread over several GPUs).  Frames that need composition
// (crops, blending, reference slots) report JXLB_UNSUPPORTED.
#include <cmath>
#include <cstdlib>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/jxlb200.h"
#include "decoder.h"
#include "frame_parser.h"

using namespace jxlb;

struct jxlb_anim {
  ByteVec cs;
  size_t cs_len = 0;
  ImageMetadata md;
  std::vector<FrameHeader> frames;       // displayed frames (regular / skip-progressive)
  std::vector<int32_t> durations_ms;
  int32_t cfg = 1, scale_mode = 1, filter = 1, api_level = 34;
  // Frames are independent pictures and a single small frame leaves the GPU almost idle (its entropy stage is a handful
  // of serial chains), so getFrame(i) decodes frames i .. i + prefetch - 1 in ONE batch and keeps the others for the
  // calls that follow -- players ask for the frames in order.  Each request is a mini codestream: the image header
  // followed by that frame's bytes only (frames are byte-aligned).
  size_t header_bytes = 0;
  std::vector<std::pair<size_t, size_t>> frame_bytes;   // [begin, end) of every displayed frame
  int prefetch = 32;   // capped so that the prefetched pictures stay below ~256 MB (jxlb_anim_open)
  std::mutex mu;
  struct Key {
    int32_t frame, w, h;
    bool operator<(const Key& o) const { return frame != o.frame ? frame < o.frame : w != o.w ? w < o.w : h < o.h; }
  };
  std::map<Key, DecodedImage> cache;
  std::deque<Key> cache_order;
  // Read-ahead: when a prefetched block is handed out, the NEXT block is submitted at once (SubmitBatch: device half on a
  // worker thread), so that a player walking through the frames finds it decoded -- a block alone is a latency chain
  // (its LF stage) during which the GPU is almost idle, and two blocks in flight overlap.
  PendingBatch* ahead = nullptr;
  int32_t ahead_first = -1, ahead_n = 0, ahead_w = 0, ahead_h = 0;
  void DropAhead() {
    if (!ahead) return;
    std::vector<DecodedImage> res;
    CollectBatch(ahead, &res, nullptr);
    for (auto& d : res) FreeImageMemory(d.data, d.device);
    ahead = nullptr;
    ahead_first = -1;
  }
  ~jxlb_anim() {
    DropAhead();
    for (auto& kv : cache) FreeImageMemory(kv.second.data, kv.second.device);
  }
};

extern "C" {

static jxlb_anim* AnimOpen(const uint8_t* data, size_t len, int32_t color_config, int32_t scale_mode, int32_t filter,
                           int32_t api_level, int32_t* status);
static int AnimGetFrame(jxlb_anim* a, int32_t frame, int32_t width, int32_t height, jxlb_image* out);

// nothing may unwind across the C boundary: allocation failures become JXLB_OOM (the reference's "not enough memory")
jxlb_anim* jxlb_anim_open(const uint8_t* data, size_t len, int32_t color_config, int32_t scale_mode, int32_t filter,
                          int32_t api_level, int32_t* status) {
  try {
    return AnimOpen(data, len, color_config, scale_mode, filter, api_level, status);
  } catch (...) {
    if (status) *status = JXLB_OOM;
    return nullptr;
  }
}
int jxlb_anim_get_frame(jxlb_anim* a, int32_t frame, int32_t width, int32_t height, jxlb_image* out) {
  try {
    return AnimGetFrame(a, frame, width, height, out);
  } catch (...) {
    if (out) {
      memset(out, 0, sizeof *out);
      out->device = -1;
      snprintf(out->message, sizeof out->message, "Not enough memory to decode this image");
    }
    return JXLB_OOM;
  }
}

static jxlb_anim* AnimOpen(const uint8_t* data, size_t len, int32_t color_config, int32_t scale_mode, int32_t filter,
                           int32_t api_level, int32_t* status) {
  auto set = [&](int s) {
    if (status) *status = s;
  };
  if (color_config < 1 || color_config > 6 || scale_mode < 1 || scale_mode > 3 || filter < 1 || filter > 10) {
    set(JXLB_BAD_ARG);
    return nullptr;
  }
  std::unique_ptr<jxlb_anim> holder(new jxlb_anim());
  jxlb_anim* a = holder.get();
  a->cfg = color_config;
  a->scale_mode = scale_mode;
  a->filter = filter;
  a->api_level = api_level > 0 ? api_level : 34;
  std::string err;
  uint64_t fb = 0;
  if (!data || ExtractCodestream(data, len, &a->cs, &a->cs_len) ||
      ParseImageHeader(a->cs.data(), a->cs.size(), a->cs_len, &a->md, &fb, &err)) {
    set(JXLB_INVALID_JXL);
    return nullptr;
  }
  a->header_bytes = (size_t) ((fb + 7) / 8);
  const bool byte_aligned_header = fb % 8 == 0;
  {
    const uint64_t frame_out = (uint64_t) a->md.xsize * a->md.ysize * 4;
    a->prefetch = (int) std::min<uint64_t>(32, std::max<uint64_t>(1, (256ull << 20) / std::max<uint64_t>(frame_out, 1)));
  }
  if (const char* e = getenv("JXLB_ANIM_PREFETCH")) a->prefetch = std::max(1, atoi(e));
  bool saved[4] = {false, false, false, false};  // reference slots written so far (same rule as ParseRequest, decoder.cu)
  for (;;) {
    FrameHeader fh;
    const uint64_t frame_begin_bit = fb;
    if (ParseFrameHeader(a->cs.data(), a->cs.size(), a->cs_len, a->md, fb, &fh, &err)) {
      set(JXLB_INVALID_JXL);
      return nullptr;
    }
    if (fh.frame_type == 0 || fh.frame_type == 3) {
      // duration in ms = round(1000 * duration * tps_den / tps_num) in float (JxlAnimatedDecoder.hpp:142-156)
      float ms = 0.f;
      if (a->md.have_animation && a->md.tps_num)
        ms = roundf(1000.0f * (float) fh.duration * (float) a->md.tps_den / (float) a->md.tps_num);
      a->durations_ms.push_back((int32_t) ms);
      a->frames.push_back(fh);
      a->frame_bytes.emplace_back((size_t) (frame_begin_bit / 8), (size_t) fh.end_byte);
      if (!byte_aligned_header || frame_begin_bit % 8 != 0) a->prefetch = 0;  // fall back to whole-file requests
      // a frame that needs the canvas of earlier frames is not an independent picture: the whole-file path below refuses
      // those (a cropped kReplace frame over never-written slots is independent: decoder.cu places it on a cleared canvas)
      const bool covers = !fh.have_crop && fh.coded_w == a->md.xsize && fh.coded_h == a->md.ysize;
      bool independent = fh.blend.mode == 0;
      for (const BlendingInfo& b : fh.ec_blend) independent = independent && b.mode == 0;
      if (!covers) {
        independent = independent && !saved[fh.blend.source & 3];
        for (const BlendingInfo& b : fh.ec_blend) independent = independent && !saved[b.source & 3];
      }
      if (!independent) a->prefetch = 0;
    } else {
      a->prefetch = 0;  // invisible (reference-only) frames: displayed frames are not self-contained
    }
    if (fh.is_last) break;
    if (fh.frame_type == 2 || (fh.frame_type != 1 && (fh.duration == 0 || fh.save_as_reference != 0))) saved[fh.save_as_reference & 3] = true;
    fb = fh.end_byte * 8;
  }
  set(JXLB_OK);
  return holder.release();
}

int32_t jxlb_anim_num_frames(const jxlb_anim* a) { return a ? (int32_t) a->frames.size() : 0; }
int32_t jxlb_anim_frame_duration_ms(const jxlb_anim* a, int32_t frame) {
  if (!a || frame < 0 || frame >= (int32_t) a->durations_ms.size()) return 0;
  return a->durations_ms[frame];
}
int32_t jxlb_anim_loops(const jxlb_anim* a) { return a ? (int32_t) a->md.num_loops : 0; }
int32_t jxlb_anim_width(const jxlb_anim* a) { return a ? (int32_t) (a->md.orientation >= 5 ? a->md.ysize : a->md.xsize) : 0; }
int32_t jxlb_anim_height(const jxlb_anim* a) { return a ? (int32_t) (a->md.orientation >= 5 ? a->md.xsize : a->md.ysize) : 0; }
static int AnimGetFrame(jxlb_anim* a, int32_t frame, int32_t width, int32_t height, jxlb_image* out) {
  if (!out) return JXLB_BAD_ARG;
  memset(out, 0, sizeof *out);
  out->device = -1;
  if (!a || frame < 0 || frame >= (int32_t) a->frames.size()) {
    snprintf(out->message, sizeof out->message, "frame index out of range");
    return JXLB_BAD_ARG;
  }
  // JxlAnimatedDecoderCoordinator.cpp:162-: rescale only when both target dimensions are positive
  const bool rescale = width > 0 && height > 0 && (width != jxlb_anim_width(a) || height != jxlb_anim_height(a));
  // RescaleImage with the coordinator's scale mode and sampler (same refusals as decodeSampled: resize.h)
  const int32_t rw = rescale ? width : -1, rh = rescale ? height : -1;
  DecodedImage d;
  {
    std::lock_guard<std::mutex> lock(a->mu);
    const jxlb_anim::Key key{frame, rw, rh};
    auto it = a->cache.find(key);
    if (it != a->cache.end()) {
      d = it->second;  // ownership moves to the caller
      a->cache.erase(it);
      for (auto ko = a->cache_order.begin(); ko != a->cache_order.end(); ++ko)  // no stale keys: eviction stays oldest-first
        if (!(*ko < key) && !(key < *ko)) {
          a->cache_order.erase(ko);
          break;
        }
    } else if (a->prefetch <= 0) {
      jxlb_request r{a->cs.data(), a->cs_len, rw, rh, a->cfg, a->scale_mode, a->filter};
      std::vector<DecodedImage> res;
      BatchTimings tm;
      const int32_t fi = frame;
      DecodeBatch(&r, 1, a->api_level, -1, -1, &res, &tm, &fi);
      d = res[0];
    } else {
      // requests of the block of frames [first, first + n): mini codestreams (image header + one frame each)
      auto submit_block = [&](int32_t first, int32_t n) {
        std::vector<std::vector<uint8_t>> mini(n);
        std::vector<jxlb_request> reqs(n);
        for (int32_t k = 0; k < n; ++k) {
          const auto& fb = a->frame_bytes[first + k];
          mini[k].reserve(a->header_bytes + (fb.second - fb.first));
          mini[k].assign(a->cs.begin(), a->cs.begin() + a->header_bytes);
          mini[k].insert(mini[k].end(), a->cs.begin() + fb.first, a->cs.begin() + fb.second);
          reqs[k] = jxlb_request{mini[k].data(), mini[k].size(), rw, rh, a->cfg, a->scale_mode, a->filter};
        }
        std::vector<int32_t> fidx(n, 0);  // every mini codestream holds one frame: displayed frame 0
        return SubmitBatch(reqs.data(), (size_t) n, a->api_level, -1, -1, fidx.data());
      };
      const int32_t n = std::min<int32_t>(a->prefetch, (int32_t) a->frames.size() - frame);
      std::vector<DecodedImage> res;
      BatchTimings tm;
      // One block beyond the one being waited for is always in flight: the next block is submitted BEFORE this one is
      // collected, so two blocks overlap on the GPU even when the caller asks for the frames back to back.
      PendingBatch* mine = nullptr;
      if (a->ahead && a->ahead_first == frame && a->ahead_n == n && a->ahead_w == rw && a->ahead_h == rh) {
        mine = a->ahead;  // the read-ahead block is exactly what is asked for
        a->ahead = nullptr;
        a->ahead_first = -1;
      } else {
        a->DropAhead();
        mine = submit_block(frame, n);
      }
      {
        const int32_t nf = frame + n, nn = std::min<int32_t>(a->prefetch, (int32_t) a->frames.size() - nf);
        if (nn > 0 && !a->cache.count(jxlb_anim::Key{nf, rw, rh})) {
          a->ahead = submit_block(nf, nn);
          a->ahead_first = nf;
          a->ahead_n = nn;
          a->ahead_w = rw;
          a->ahead_h = rh;
        }
      }
      CollectBatch(mine, &res, &tm);
      d = res[0];
      for (int32_t k = 1; k < n; ++k) {
        const jxlb_anim::Key kk{frame + k, rw, rh};
        if (res[k].status != JXLB_OK || a->cache.count(kk)) {
          FreeImageMemory(res[k].data, res[k].device);
          continue;
        }
        a->cache[kk] = res[k];
        a->cache_order.push_back(kk);
      }
      while (a->cache.size() > (size_t) 2 * a->prefetch && !a->cache_order.empty()) {  // bounded: drop the oldest leftovers
        auto old = a->cache.find(a->cache_order.front());
        a->cache_order.pop_front();
        if (old != a->cache.end()) {
          FreeImageMemory(old->second.data, old->second.device);
          a->cache.erase(old);
        }
      }
    }
  }
  out->data = d.data;
  out->width = d.width;
  out->height = d.height;
  out->stride_bytes = d.stride_bytes;
  out->format = d.format;
  out->color_space = d.color_space;
  out->premultiplied = d.premultiplied;
  out->device = d.device;
  snprintf(out->message, sizeof out->message, "%s", d.message.c_str());
  return d.status;
}
void jxlb_anim_close(jxlb_anim* a) { delete a; }

}  // extern "C"
