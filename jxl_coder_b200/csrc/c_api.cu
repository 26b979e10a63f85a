// extern "C" surface of libjxlb200.so (include/jxlb200.h).
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/jxlb200.h"
#include "decoder.h"
#include "frame_parser.h"
#include "kernels.h"

using namespace jxlb;

namespace {
std::mutex g_tm_mu;
BatchTimings g_last_timings;

void FillImage(const DecodedImage& d, jxlb_image* out) {
  memset(out, 0, sizeof *out);
  out->data = d.data;
  out->width = d.width;
  out->height = d.height;
  out->stride_bytes = d.stride_bytes;
  out->format = d.format;
  out->color_space = d.color_space;
  out->premultiplied = d.premultiplied;
  out->device = d.device;
  snprintf(out->message, sizeof out->message, "%s", d.message.c_str());
}
}  // namespace

namespace {
int FailAll(jxlb_image* outs, int32_t* status, size_t n, int code, const char* msg) {
  for (size_t i = 0; i < n; ++i) {
    memset(&outs[i], 0, sizeof outs[i]);
    outs[i].device = -1;
    snprintf(outs[i].message, sizeof outs[i].message, "%s", msg);
    if (status) status[i] = code;
  }
  return code;
}
}  // namespace

extern "C" {

int jxlb_decode_batch(const jxlb_request* reqs, size_t n, jxlb_image* outs, int32_t* status, const jxlb_batch_opts* opts) {
  if (!reqs || !outs) return JXLB_BAD_ARG;
  jxlb_batch_opts o{0, -1, -1, 0};
  if (opts) o = *opts;
  // nothing may unwind across the C boundary (the JNI entry points catch std::bad_alloc / std::runtime_error and throw
  // Java exceptions instead, JniDecoding.cpp:81-93)
  try {
    std::vector<DecodedImage> res;
    BatchTimings tm;
    int rc = DecodeBatch(reqs, n, o.api_level, o.device, o.output_device, &res, &tm);
    {
      std::lock_guard<std::mutex> l(g_tm_mu);
      g_last_timings = tm;
    }
    for (size_t i = 0; i < n; ++i) {
      FillImage(res[i], &outs[i]);
      if (status) status[i] = res[i].status;
    }
    return rc;
  } catch (const std::bad_alloc&) {
    return FailAll(outs, status, n, JXLB_OOM, "Not enough memory to decode this image");
  } catch (...) {
    return FailAll(outs, status, n, JXLB_ERROR, "Error while decoding");
  }
}

struct jxlb_pending {
  jxlb::PendingBatch* p;
  size_t n;
};

jxlb_pending* jxlb_decode_batch_submit(const jxlb_request* reqs, size_t n, const jxlb_batch_opts* opts) {
  if (!reqs || !n) return nullptr;
  jxlb_batch_opts o{0, -1, -1, 0};
  if (opts) o = *opts;
  try {
    jxlb::PendingBatch* p = SubmitBatch(reqs, n, o.api_level, o.device, o.output_device);
    return new jxlb_pending{p, n};
  } catch (...) {
    return nullptr;
  }
}

int jxlb_decode_batch_collect(jxlb_pending* h, jxlb_image* outs, int32_t* status) {
  if (!h || !outs) return JXLB_BAD_ARG;
  const size_t n = h->n;
  try {
    std::vector<DecodedImage> res;
    BatchTimings tm;
    const int rc = CollectBatch(h->p, &res, &tm);
    delete h;
    {
      std::lock_guard<std::mutex> l(g_tm_mu);
      g_last_timings = tm;
    }
    for (size_t i = 0; i < n; ++i) {
      FillImage(res[i], &outs[i]);
      if (status) status[i] = res[i].status;
    }
    return rc;
  } catch (const std::bad_alloc&) {
    return FailAll(outs, status, n, JXLB_OOM, "Not enough memory to decode this image");
  } catch (...) {
    return FailAll(outs, status, n, JXLB_ERROR, "Error while decoding");
  }
}

int jxlb_decode_sampled(const uint8_t* data, size_t len, int32_t width, int32_t height, int32_t color_config, int32_t scale_mode,
                        int32_t filter, int32_t api_level, jxlb_image* out) {
  if (!out) return JXLB_BAD_ARG;
  jxlb_request r{data, len, width, height, color_config, scale_mode, filter};
  jxlb_batch_opts o{api_level, -1, -1, 0};
  int32_t st = JXLB_OK;
  jxlb_decode_batch(&r, 1, out, &st, &o);
  return st;
}

int jxlb_test_recon_block(int device, uint32_t strategy, const int16_t* q, const float* lf, uint32_t hf_mul, uint32_t global_scale,
                          float* out) try {
  return TestReconBlock(device, strategy, q, lf, hf_mul, global_scale, out);
} catch (...) {
  return JXLB_ERROR;
}

int jxlb_get_size(const uint8_t* data, size_t len, uint32_t* width, uint32_t* height) try {
  // DecodeBasicInfo (interop/JxlDecoding.cpp:178-226): header-only, CPU.
  if (!data || !width || !height) return JXLB_BAD_ARG;
  ByteVec cs;
  size_t cs_len = 0;
  int st = ExtractCodestream(data, len, &cs, &cs_len);
  if (st == kParseNotJxl) return JXLB_NOT_JXL;
  if (st) return JXLB_INVALID_JXL;
  ImageMetadata md;
  uint64_t fb = 0;
  std::string err;
  st = ParseImageHeader(cs.data(), cs.size(), cs_len, &md, &fb, &err);
  if (st == kParseNotJxl) return JXLB_NOT_JXL;
  if (st == kParseInvalid) return JXLB_INVALID_JXL;
  // an unsupported feature later in the headers does not prevent reporting the size
  if (md.xsize == 0 || md.ysize == 0) return JXLB_INVALID_JXL;
  *width = md.xsize;
  *height = md.ysize;
  if (md.orientation >= 5) {
    *width = md.ysize;
    *height = md.xsize;
  }
  return JXLB_OK;
} catch (...) {
  return JXLB_OOM;
}

void jxlb_image_free(jxlb_image* img) {
  if (!img || !img->data) return;
  FreeImageMemory(img->data, img->device);
  img->data = nullptr;
}

// ---- prepared batches ----
struct jxlb_batch {
  jxlb::Batch* b;
};

jxlb_batch* jxlb_batch_prepare(const jxlb_request* reqs, size_t n, const jxlb_batch_opts* opts, int32_t* status) {
  if (!reqs || !n) return nullptr;
  jxlb_batch_opts o{0, -1, -1, 0};
  if (opts) o = *opts;
  try {
    std::vector<int> st;
    jxlb::Batch* b = PrepareBatch(reqs, n, o.api_level, o.device, &st);
    if (status)
      for (size_t i = 0; i < n; ++i) status[i] = st[i];
    jxlb_batch* h = new jxlb_batch{b};
    return h;
  } catch (...) {
    if (status)
      for (size_t i = 0; i < n; ++i) status[i] = JXLB_OOM;
    return nullptr;
  }
}
int jxlb_batch_run(jxlb_batch* b) { return b ? RunBatch(b->b, true) : JXLB_BAD_ARG; }
int jxlb_batch_run_async(jxlb_batch* b) { return b ? RunBatch(b->b, false) : JXLB_BAD_ARG; }
int jxlb_batch_wait(jxlb_batch* b) { return b ? WaitBatch(b->b) : JXLB_BAD_ARG; }
int jxlb_batch_fetch(jxlb_batch* b, size_t index, jxlb_image* out) {
  if (!b || !out) return JXLB_BAD_ARG;
  DecodedImage d;
  int rc = FetchBatchImage(b->b, index, &d);
  FillImage(d, out);
  return rc;
}
const void* jxlb_batch_device_pixels(const jxlb_batch* b, size_t index, size_t* bytes) {
  return b ? BatchDevicePixels(b->b, index, bytes) : nullptr;
}
void jxlb_batch_stage_ms(const jxlb_batch* b, float* ms8) { BatchStageMs(b ? b->b : nullptr, ms8); }
void jxlb_batch_stage_ms_mean(const jxlb_batch* b, float* ms8, int32_t* runs) {
  int r = 0;
  BatchStageMsMean(b ? b->b : nullptr, ms8, &r);
  if (runs) *runs = r;
}
void jxlb_batch_reset_stats(jxlb_batch* b) {
  if (b) ResetBatchStats(b->b);
}
float jxlb_batch_span_ms(const jxlb_batch* first, const jxlb_batch* last) {
  return (first && last) ? BatchSpanMs(first->b, last->b) : -1.f;
}
void jxlb_batch_free(jxlb_batch* b) {
  if (!b) return;
  FreeBatch(b->b);
  delete b;
}

// ---- animated images: implemented in anim.cu ----

uint64_t jxlb_kernel_launches(void) { return KernelLaunchCount(); }
void jxlb_last_batch_timings(float* ms6) {
  std::lock_guard<std::mutex> l(g_tm_mu);
  for (int i = 0; i < 6; ++i) ms6[i] = g_last_timings.ms[i];
}
const char* jxlb_version(void) { return "jxlb200 0.1 (sm_100a)"; }

}  // extern "C"
