// JxlAnimatedImage entry points (include/jxlb200.h).  Round-1 state: the frame table (count, durations, loops, size)
// is parsed on the CPU exactly as the reference's JxlAnimatedDecoder constructor does
// (/root/reference/jxlcoder/src/main/cpp/interop/JxlAnimatedDecoder.hpp:68-185); per-frame pixel decode with
// blending / reference slots is not implemented yet and reports JXLB_UNSUPPORTED.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/jxlb200.h"
#include "frame_parser.h"

using namespace jxlb;

struct jxlb_anim {
  std::vector<uint8_t> cs;
  size_t cs_len = 0;
  ImageMetadata md;
  std::vector<FrameHeader> frames;       // displayed frames (regular / skip-progressive)
  std::vector<int32_t> durations_ms;
  int32_t cfg = 1, scale_mode = 1, filter = 1, api_level = 34;
};

extern "C" {

jxlb_anim* jxlb_anim_open(const uint8_t* data, size_t len, int32_t color_config, int32_t scale_mode, int32_t filter,
                          int32_t api_level, int32_t* status) {
  auto set = [&](int s) {
    if (status) *status = s;
  };
  if (color_config < 1 || color_config > 6 || scale_mode < 1 || scale_mode > 3 || filter < 1 || filter > 10) {
    set(JXLB_BAD_ARG);
    return nullptr;
  }
  jxlb_anim* a = new jxlb_anim();
  a->cfg = color_config;
  a->scale_mode = scale_mode;
  a->filter = filter;
  a->api_level = api_level > 0 ? api_level : 34;
  std::string err;
  uint64_t fb = 0;
  if (!data || ExtractCodestream(data, len, &a->cs, &a->cs_len) ||
      ParseImageHeader(a->cs.data(), a->cs.size(), a->cs_len, &a->md, &fb, &err)) {
    delete a;
    set(JXLB_INVALID_JXL);
    return nullptr;
  }
  for (;;) {
    FrameHeader fh;
    if (ParseFrameHeader(a->cs.data(), a->cs.size(), a->cs_len, a->md, fb, &fh, &err)) {
      delete a;
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
    }
    if (fh.is_last) break;
    fb = fh.end_byte * 8;
  }
  set(JXLB_OK);
  return a;
}

int32_t jxlb_anim_num_frames(const jxlb_anim* a) { return a ? (int32_t) a->frames.size() : 0; }
int32_t jxlb_anim_frame_duration_ms(const jxlb_anim* a, int32_t frame) {
  if (!a || frame < 0 || frame >= (int32_t) a->durations_ms.size()) return 0;
  return a->durations_ms[frame];
}
int32_t jxlb_anim_loops(const jxlb_anim* a) { return a ? (int32_t) a->md.num_loops : 0; }
int32_t jxlb_anim_width(const jxlb_anim* a) { return a ? (int32_t) a->md.xsize : 0; }
int32_t jxlb_anim_height(const jxlb_anim* a) { return a ? (int32_t) a->md.ysize : 0; }
int jxlb_anim_get_frame(jxlb_anim* a, int32_t frame, int32_t width, int32_t height, jxlb_image* out) {
  (void) width;
  (void) height;
  if (out) {
    memset(out, 0, sizeof *out);
    snprintf(out->message, sizeof out->message, "animated frame decode is not implemented yet");
  }
  if (!a || frame < 0 || frame >= (int32_t) a->frames.size()) return JXLB_BAD_ARG;
  return JXLB_UNSUPPORTED;
}
void jxlb_anim_close(jxlb_anim* a) { delete a; }

}  // extern "C"
