// TEST INFRASTRUCTURE (oracle/_ref) — never linked into the product, never measured as the product.
//
// C entry points around the reference's OWN code so that Python tests / bench.py can call it with ctypes:
//   * ref_decode_sampled      → Java_com_awxkee_jxlcoder_JxlCoder_decodeSampledImpl / decodeByteBufferSampledImpl
//                               (/root/reference/jxlcoder/src/main/cpp/JniDecoding.cpp:333-392), unmodified, through the
//                               fake JNI in fakejni/.
//   * ref_get_size            → Java_com_awxkee_jxlcoder_JxlCoder_getSizeImpl (JniDecoding.cpp:394-414)
//   * ref_anim_*              → Java_com_awxkee_jxlcoder_JxlAnimatedImage_* (JxlAnimatedDecoderCoordinator.cpp:45-425)
//   * ref_decode_oneshot      → DecodeJpegXlOneShot (interop/JxlDecoding.cpp:36-176): raw libjxl output before post-process
//   * ref_encode / ref_anim_encode → EncodeJxlOneshot (interop/JxlEncoding.cpp:48-193) / JxlAnimatedEncoder
//                               (interop/JxlAnimatedEncoder.hpp): used ONLY to generate synthetic inputs.
//   * ref_encode_ex           → libjxl encoder API directly with extra frame settings (EPF/Gaborish/effort/...) to
//                               stage coding tools for tests.
//   * ref_pack_* / ref_weave_u8/u16 → thin C wrappers over imagebit/*.cpp and the prebuilt libweaver.a.
#include <jni.h>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "fakejni.h"
#include "interop/JxlDecoding.h"
#include "interop/JxlEncoding.h"
#include "interop/JxlAnimatedEncoder.hpp"
#include "imagebit/RGBAlpha.h"
#include "imagebit/Rgba8ToF16.h"
#include "imagebit/RgbaU16toHF.h"
#include "imagebit/Rgb1010102.h"
#include "imagebit/Rgb565.h"
#include "imagebit/Rgba16.h"
#include "weaver.h"
#include "jxl/encode.h"
#include "jxl/encode_cxx.h"
#include "jxl/decode.h"
#include "jxl/thread_parallel_runner.h"
#include "jxl/thread_parallel_runner_cxx.h"

extern "C" {
// The reference's JNI exports (defined in its own translation units).
jobject Java_com_awxkee_jxlcoder_JxlCoder_decodeSampledImpl(JNIEnv *, jobject, jbyteArray, jint, jint, jint, jint, jint);
jobject Java_com_awxkee_jxlcoder_JxlCoder_decodeByteBufferSampledImpl(JNIEnv *, jobject, jobject, jint, jint, jint, jint, jint);
jobject Java_com_awxkee_jxlcoder_JxlCoder_getSizeImpl(JNIEnv *, jobject, jbyteArray);
jlong Java_com_awxkee_jxlcoder_JxlAnimatedImage_createCoordinator(JNIEnv *, jobject, jobject, jint, jint, jint);
jlong Java_com_awxkee_jxlcoder_JxlAnimatedImage_createCoordinatorByteArray(JNIEnv *, jobject, jbyteArray, jint, jint, jint);
void Java_com_awxkee_jxlcoder_JxlAnimatedImage_closeAndReleaseAnimatedImage(JNIEnv *, jobject, jlong);
jint Java_com_awxkee_jxlcoder_JxlAnimatedImage_getNumberOfFrames(JNIEnv *, jobject, jlong);
jint Java_com_awxkee_jxlcoder_JxlAnimatedImage_getFrameDurationImpl(JNIEnv *, jobject, jlong, jint);
jint Java_com_awxkee_jxlcoder_JxlAnimatedImage_getLoopsCount(JNIEnv *, jobject, jlong);
jint Java_com_awxkee_jxlcoder_JxlAnimatedImage_getHeightImpl(JNIEnv *, jobject, jlong);
jint Java_com_awxkee_jxlcoder_JxlAnimatedImage_getWidthImpl(JNIEnv *, jobject, jlong);
jobject Java_com_awxkee_jxlcoder_JxlAnimatedImage_getFrameImpl(JNIEnv *, jobject, jlong, jint, jint, jint);
}

struct ref_image {
  uint8_t *data;       // malloc'ed copy of the Bitmap pixels (ref_free)
  uint32_t width, height, stride;
  char config[24];       // Bitmap.Config name: ARGB_8888 | RGBA_F16 | RGB_565 | RGBA_1010102
  char color_space[24];  // ColorSpace.Named passed to createBitmap ("" when api_level < 34)
  char error_class[80];  // Java exception class raised ("" = none)
  char error_msg[256];
};

namespace {
struct Env {
  FakeEnvState state;
  JNIEnv env;
  Env() { env.impl = &state; }
};

int finish(Env &e, jobject bmp, ref_image *out) {
  memset(out, 0, sizeof(*out));
  if (e.state.has_exception || !bmp) {
    snprintf(out->error_class, sizeof out->error_class, "%s", e.state.exception_class.c_str());
    snprintf(out->error_msg, sizeof out->error_msg, "%s", e.state.exception_msg.c_str());
    return 1;
  }
  out->width = (uint32_t) bmp->width;
  out->height = (uint32_t) bmp->height;
  out->stride = (uint32_t) bmp->stride;
  snprintf(out->config, sizeof out->config, "%s", bmp->config.c_str());
  snprintf(out->color_space, sizeof out->color_space, "%s", bmp->colorspace.c_str());
  out->data = (uint8_t *) malloc(bmp->pixels.size() ? bmp->pixels.size() : 1);
  memcpy(out->data, bmp->pixels.data(), bmp->pixels.size());
  return 0;
}
}  // namespace

extern "C" {

void ref_set_api_level(int level) { fakejni_set_api_level(level); }
void ref_free(void *p) { free(p); }

// buffer_kind: 0 = jbyteArray entry point, 1 = direct ByteBuffer, 2 = non-direct ByteBuffer (must raise).
int ref_decode_sampled(const uint8_t *data, size_t len, int w, int h, int cfg, int scale_mode, int filter,
                       int buffer_kind, ref_image *out) {
  Env e;
  jobject bmp;
  if (buffer_kind == 0) {
    auto *arr = e.state.make(FakeObject::ByteArray, "byte[]");
    arr->data = data;
    arr->len = len;
    bmp = Java_com_awxkee_jxlcoder_JxlCoder_decodeSampledImpl(&e.env, nullptr, arr, w, h, cfg, scale_mode, filter);
  } else {
    auto *buf = e.state.make(buffer_kind == 1 ? FakeObject::DirectBuffer : FakeObject::HeapBuffer, "ByteBuffer");
    buf->data = data;
    buf->len = len;
    bmp = Java_com_awxkee_jxlcoder_JxlCoder_decodeByteBufferSampledImpl(&e.env, nullptr, buf, w, h, cfg, scale_mode, filter);
  }
  return finish(e, bmp, out);
}

// returns 1 and fills w,h; 0 when the reference returns null.
int ref_get_size(const uint8_t *data, size_t len, uint32_t *w, uint32_t *h) {
  Env e;
  auto *arr = e.state.make(FakeObject::ByteArray, "byte[]");
  arr->data = data;
  arr->len = len;
  jobject s = Java_com_awxkee_jxlcoder_JxlCoder_getSizeImpl(&e.env, nullptr, arr);
  if (!s) return 0;
  *w = (uint32_t) s->width;
  *h = (uint32_t) s->height;
  return 1;
}

// ---- animated images ---------------------------------------------------------------------------------------------
int64_t ref_anim_open(const uint8_t *data, size_t len, int cfg, int scale_mode, int filter, int use_bytebuffer,
                      char *err, size_t err_len) {
  Env e;
  jlong handle;
  if (use_bytebuffer) {
    auto *buf = e.state.make(FakeObject::DirectBuffer, "ByteBuffer");
    buf->data = data;
    buf->len = len;
    handle = Java_com_awxkee_jxlcoder_JxlAnimatedImage_createCoordinator(&e.env, nullptr, buf, cfg, scale_mode, filter);
  } else {
    auto *arr = e.state.make(FakeObject::ByteArray, "byte[]");
    arr->data = data;
    arr->len = len;
    handle = Java_com_awxkee_jxlcoder_JxlAnimatedImage_createCoordinatorByteArray(&e.env, nullptr, arr, cfg, scale_mode, filter);
  }
  if (err && err_len) snprintf(err, err_len, "%s: %s", e.state.exception_class.c_str(), e.state.exception_msg.c_str());
  return handle;
}
void ref_anim_close(int64_t hnd) { Env e; Java_com_awxkee_jxlcoder_JxlAnimatedImage_closeAndReleaseAnimatedImage(&e.env, nullptr, hnd); }
int ref_anim_num_frames(int64_t hnd) { Env e; return Java_com_awxkee_jxlcoder_JxlAnimatedImage_getNumberOfFrames(&e.env, nullptr, hnd); }
int ref_anim_frame_duration(int64_t hnd, int i) { Env e; return Java_com_awxkee_jxlcoder_JxlAnimatedImage_getFrameDurationImpl(&e.env, nullptr, hnd, i); }
int ref_anim_loops(int64_t hnd) { Env e; return Java_com_awxkee_jxlcoder_JxlAnimatedImage_getLoopsCount(&e.env, nullptr, hnd); }
int ref_anim_width(int64_t hnd) { Env e; return Java_com_awxkee_jxlcoder_JxlAnimatedImage_getWidthImpl(&e.env, nullptr, hnd); }
int ref_anim_height(int64_t hnd) { Env e; return Java_com_awxkee_jxlcoder_JxlAnimatedImage_getHeightImpl(&e.env, nullptr, hnd); }
int ref_anim_get_frame(int64_t hnd, int frame, int w, int h, ref_image *out) {
  Env e;
  jobject bmp = Java_com_awxkee_jxlcoder_JxlAnimatedImage_getFrameImpl(&e.env, nullptr, hnd, frame, w, h);
  return finish(e, bmp, out);
}

// ---- raw libjxl output (what the GPU decode stage must reproduce) -------------------------------------------------
struct ref_raw {
  uint8_t *pixels;  // malloc'ed, RGBA interleaved u8 or u16
  size_t xsize, ysize;
  uint32_t bit_depth;  // 8 or 16 (as DecodeJpegXlOneShot reports)
  int use_floats, alpha_premultiplied, orientation, prefer_encoding, has_alpha;
  float intensity_target;
  int color_space, white_point, primaries, transfer_function;
  double gamma;
  size_t icc_size;
};
int ref_decode_oneshot(const uint8_t *data, size_t len, int allowed_floats, ref_raw *out) {
  memset(out, 0, sizeof *out);
  std::vector<uint8_t> pixels, icc;
  bool useFloats = false, premul = false, prefer = false, hasAlpha = true;
  uint32_t depth = 8;
  JxlOrientation orient = JXL_ORIENT_IDENTITY;
  JxlColorEncoding ce;
  memset(&ce, 0, sizeof ce);
  float it = 255.f;
  try {
    if (!DecodeJpegXlOneShot(data, len, &pixels, &out->xsize, &out->ysize, &icc, &useFloats, &depth, &premul,
                             allowed_floats != 0, &orient, &prefer, &ce, &hasAlpha, &it))
      return 1;
  } catch (InvalidImageSizeException &) {
    return 2;
  } catch (std::exception &) {
    return 3;
  }
  out->pixels = (uint8_t *) malloc(pixels.size() ? pixels.size() : 1);
  memcpy(out->pixels, pixels.data(), pixels.size());
  out->bit_depth = depth;
  out->use_floats = useFloats;
  out->alpha_premultiplied = premul;
  out->orientation = (int) orient;
  out->prefer_encoding = prefer;
  out->has_alpha = hasAlpha;
  out->intensity_target = it;
  out->color_space = ce.color_space;
  out->white_point = ce.white_point;
  out->primaries = ce.primaries;
  out->transfer_function = ce.transfer_function;
  out->gamma = ce.gamma;
  out->icc_size = icc.size();
  return 0;
}

// Time-only variant for the CPU baseline: decodes into a caller-provided scratch (no malloc/copy of the result
// beyond what DecodeJpegXlOneShot itself does). Returns 0 on success.
int ref_decode_oneshot_discard(const uint8_t *data, size_t len, int allowed_floats) {
  std::vector<uint8_t> pixels, icc;
  size_t xs = 0, ys = 0;
  bool useFloats = false, premul = false, prefer = false, hasAlpha = true;
  uint32_t depth = 8;
  JxlOrientation orient = JXL_ORIENT_IDENTITY;
  JxlColorEncoding ce;
  float it = 255.f;
  try {
    return DecodeJpegXlOneShot(data, len, &pixels, &xs, &ys, &icc, &useFloats, &depth, &premul, allowed_floats != 0,
                               &orient, &prefer, &ce, &hasAlpha, &it) ? 0 : 1;
  } catch (...) {
    return 2;
  }
}

// ---- encoders (input generation only) ------------------------------------------------------------------------------
// colorspace: 1 rgb, 2 rgba, 3 mono; compression: 1 lossless, 2 lossy; data_format: 1 u8, 2 u16
// primaries/transfer: JxlPrimaries / JxlTransferFunction enum ints (1/13 = sRGB)
int ref_encode(const uint8_t *pixels, size_t npix_bytes, uint32_t w, uint32_t h, int colorspace, int compression,
               int data_format, int effort, int quality, int decoding_speed, int primaries, int transfer,
               const uint8_t *icc, size_t icc_len, uint8_t **out, size_t *out_len) {
  std::vector<uint8_t> px(pixels, pixels + npix_bytes), comp, iccv;
  if (icc && icc_len) iccv.assign(icc, icc + icc_len);
  JxlColorEncoding ce;
  JxlColorEncodingSetToSRGB(&ce, colorspace == 3 ? JXL_TRUE : JXL_FALSE);
  ce.primaries = (JxlPrimaries) primaries;
  ce.transfer_function = (JxlTransferFunction) transfer;
  if (!EncodeJxlOneshot(px, w, h, &comp, (JxlColorPixelType) colorspace, (JxlCompressionOption) compression,
                        (JxlEncodingPixelDataFormat) data_format, iccv, effort, quality, decoding_speed, ce))
    return 1;
  *out = (uint8_t *) malloc(comp.size());
  memcpy(*out, comp.data(), comp.size());
  *out_len = comp.size();
  return 0;
}

// Direct libjxl encode with extra frame-setting options (ids per jxl/encode.h:132-248) and optional extra-channel
// distance; pixel layout RGB or RGBA u8/u16.
int ref_encode_ex(const uint8_t *pixels, size_t npix_bytes, uint32_t w, uint32_t h, int channels, int bits,
                  int lossless, float distance, float alpha_distance, const int *opt_ids, const int *opt_vals,
                  int nopts, int primaries, int transfer, int orientation, uint8_t **out, size_t *out_len) {
  auto enc = JxlEncoderMake(nullptr);
  auto runner = JxlThreadParallelRunnerMake(nullptr, JxlThreadParallelRunnerDefaultNumWorkerThreads());
  if (JXL_ENC_SUCCESS != JxlEncoderSetParallelRunner(enc.get(), JxlThreadParallelRunner, runner.get())) return 1;
  JxlBasicInfo bi;
  JxlEncoderInitBasicInfo(&bi);
  bi.xsize = w;
  bi.ysize = h;
  bi.bits_per_sample = bits;
  bi.uses_original_profile = lossless ? JXL_TRUE : JXL_FALSE;
  int color_channels = channels >= 3 ? 3 : 1;
  bi.num_color_channels = color_channels;
  bi.orientation = (JxlOrientation) (orientation ? orientation : 1);
  bool alpha = channels == 4 || channels == 2;
  if (alpha) {
    bi.num_extra_channels = 1;
    bi.alpha_bits = bits;
  }
  if (JXL_ENC_SUCCESS != JxlEncoderSetBasicInfo(enc.get(), &bi)) return 2;
  if (alpha) {
    JxlExtraChannelInfo ci;
    JxlEncoderInitExtraChannelInfo(JXL_CHANNEL_ALPHA, &ci);
    ci.bits_per_sample = bits;
    ci.alpha_premultiplied = 0;
    if (JXL_ENC_SUCCESS != JxlEncoderSetExtraChannelInfo(enc.get(), 0, &ci)) return 3;
  }
  JxlColorEncoding ce;
  JxlColorEncodingSetToSRGB(&ce, color_channels == 1 ? JXL_TRUE : JXL_FALSE);
  if (primaries) ce.primaries = (JxlPrimaries) primaries;
  if (transfer) ce.transfer_function = (JxlTransferFunction) transfer;
  if (JXL_ENC_SUCCESS != JxlEncoderSetColorEncoding(enc.get(), &ce)) return 4;
  JxlEncoderFrameSettings *fs = JxlEncoderFrameSettingsCreate(enc.get(), nullptr);
  if (lossless) {
    if (JXL_ENC_SUCCESS != JxlEncoderSetFrameLossless(fs, JXL_TRUE)) return 5;
  } else {
    if (JXL_ENC_SUCCESS != JxlEncoderSetFrameDistance(fs, distance)) return 5;
    if (alpha && alpha_distance >= 0 && JXL_ENC_SUCCESS != JxlEncoderSetExtraChannelDistance(fs, 0, alpha_distance)) return 6;
  }
  for (int i = 0; i < nopts; i++)
    if (JXL_ENC_SUCCESS != JxlEncoderFrameSettingsSetOption(fs, (JxlEncoderFrameSettingId) opt_ids[i], opt_vals[i])) return 7;
  JxlPixelFormat pf = {(uint32_t) channels, bits > 8 ? JXL_TYPE_UINT16 : JXL_TYPE_UINT8, JXL_NATIVE_ENDIAN, 0};
  if (JXL_ENC_SUCCESS != JxlEncoderAddImageFrame(fs, &pf, pixels, npix_bytes)) return 8;
  JxlEncoderCloseInput(enc.get());
  std::vector<uint8_t> comp(1 << 16);
  uint8_t *next = comp.data();
  size_t avail = comp.size();
  JxlEncoderStatus r = JXL_ENC_NEED_MORE_OUTPUT;
  while (r == JXL_ENC_NEED_MORE_OUTPUT) {
    r = JxlEncoderProcessOutput(enc.get(), &next, &avail);
    if (r == JXL_ENC_NEED_MORE_OUTPUT) {
      size_t off = next - comp.data();
      comp.resize(comp.size() * 2);
      next = comp.data() + off;
      avail = comp.size() - off;
    }
  }
  if (r != JXL_ENC_SUCCESS) return 9;
  size_t n = next - comp.data();
  *out = (uint8_t *) malloc(n);
  memcpy(*out, comp.data(), n);
  *out_len = n;
  return 0;
}

// Layered / composed images, libjxl encoder API directly: nlayers frames, frame i = a crop of size dims[i] = {x0, y0, w, h}
// of RGBA (or RGB) u8 pixels with blend[i] = {mode, source, save_as_reference, duration}; canvas w x h.  Used ONLY to
// stage inputs for the composition tests (cropped frames, kBlend / kAdd / kMul over reference slots).
int ref_encode_layers(const uint8_t *pixels, uint32_t w, uint32_t h, int channels, int lossless, float distance, int nlayers,
                      const int32_t *dims, const int32_t *blend, int animation, int effort, uint8_t **out, size_t *out_len) {
  auto enc = JxlEncoderMake(nullptr);
  auto runner = JxlThreadParallelRunnerMake(nullptr, JxlThreadParallelRunnerDefaultNumWorkerThreads());
  if (JXL_ENC_SUCCESS != JxlEncoderSetParallelRunner(enc.get(), JxlThreadParallelRunner, runner.get())) return 1;
  JxlBasicInfo bi;
  JxlEncoderInitBasicInfo(&bi);
  bi.xsize = w;
  bi.ysize = h;
  bi.bits_per_sample = 8;
  bi.uses_original_profile = lossless ? JXL_TRUE : JXL_FALSE;
  bi.num_color_channels = 3;
  const bool alpha = channels == 4;
  if (alpha) {
    bi.num_extra_channels = 1;
    bi.alpha_bits = 8;
  }
  if (animation) {
    bi.have_animation = JXL_TRUE;
    bi.animation.tps_numerator = 1000;
    bi.animation.tps_denominator = 1;
    bi.animation.num_loops = 0;
    bi.animation.have_timecodes = JXL_FALSE;
  }
  if (JXL_ENC_SUCCESS != JxlEncoderSetBasicInfo(enc.get(), &bi)) return 2;
  if (alpha) {
    JxlExtraChannelInfo ci;
    JxlEncoderInitExtraChannelInfo(JXL_CHANNEL_ALPHA, &ci);
    ci.bits_per_sample = 8;
    ci.alpha_premultiplied = 0;
    if (JXL_ENC_SUCCESS != JxlEncoderSetExtraChannelInfo(enc.get(), 0, &ci)) return 3;
  }
  JxlColorEncoding ce;
  JxlColorEncodingSetToSRGB(&ce, JXL_FALSE);
  if (JXL_ENC_SUCCESS != JxlEncoderSetColorEncoding(enc.get(), &ce)) return 4;
  const uint8_t *px = pixels;
  for (int i = 0; i < nlayers; ++i) {
    JxlEncoderFrameSettings *fs = JxlEncoderFrameSettingsCreate(enc.get(), nullptr);
    if (lossless) {
      if (JXL_ENC_SUCCESS != JxlEncoderSetFrameLossless(fs, JXL_TRUE)) return 5;
    } else if (JXL_ENC_SUCCESS != JxlEncoderSetFrameDistance(fs, distance)) {
      return 5;
    }
    JxlEncoderFrameSettingsSetOption(fs, JXL_ENC_FRAME_SETTING_EFFORT, effort);
    JxlFrameHeader fh;
    JxlEncoderInitFrameHeader(&fh);
    const int32_t *d = dims + 4 * i, *b = blend + 4 * i;
    fh.duration = (uint32_t) b[3];
    fh.layer_info.have_crop = JXL_TRUE;
    fh.layer_info.crop_x0 = d[0];
    fh.layer_info.crop_y0 = d[1];
    fh.layer_info.xsize = (uint32_t) d[2];
    fh.layer_info.ysize = (uint32_t) d[3];
    fh.layer_info.blend_info.blendmode = (JxlBlendMode) b[0];
    fh.layer_info.blend_info.source = (uint32_t) b[1];
    fh.layer_info.blend_info.alpha = 0;
    fh.layer_info.blend_info.clamp = JXL_FALSE;
    fh.layer_info.save_as_reference = (uint32_t) b[2];
    if (JXL_ENC_SUCCESS != JxlEncoderSetFrameHeader(fs, &fh)) return 6;
    if (alpha) {
      JxlBlendInfo abi = fh.layer_info.blend_info;
      if (JXL_ENC_SUCCESS != JxlEncoderSetExtraChannelBlendInfo(fs, 0, &abi)) return 7;
    }
    JxlPixelFormat pf = {(uint32_t) channels, JXL_TYPE_UINT8, JXL_NATIVE_ENDIAN, 0};
    const size_t bytes = (size_t) d[2] * d[3] * channels;
    if (JXL_ENC_SUCCESS != JxlEncoderAddImageFrame(fs, &pf, px, bytes)) return 8;
    px += bytes;
  }
  JxlEncoderCloseInput(enc.get());
  std::vector<uint8_t> comp(1 << 16);
  uint8_t *next = comp.data();
  size_t avail = comp.size();
  JxlEncoderStatus r = JXL_ENC_NEED_MORE_OUTPUT;
  while (r == JXL_ENC_NEED_MORE_OUTPUT) {
    r = JxlEncoderProcessOutput(enc.get(), &next, &avail);
    if (r == JXL_ENC_NEED_MORE_OUTPUT) {
      size_t off = next - comp.data();
      comp.resize(comp.size() * 2);
      next = comp.data() + off;
      avail = comp.size() - off;
    }
  }
  if (r != JXL_ENC_SUCCESS) return 9;
  size_t n = next - comp.data();
  *out = (uint8_t *) malloc(n);
  memcpy(*out, comp.data(), n);
  *out_len = n;
  return 0;
}

// Animated: nframes frames of w*h*channels u8, each with `duration` ticks (tps 1000/1), via the reference's
// JxlAnimatedEncoder (interop/JxlAnimatedEncoder.hpp:55-190).
int ref_anim_encode(const uint8_t *frames, uint32_t w, uint32_t h, int colorspace, int compression, int nframes,
                    int duration, int num_loops, int quality, int effort, int decoding_speed, uint8_t **out,
                    size_t *out_len) {
  try {
    JxlAnimatedEncoder enc((int) w, (int) h, (JxlColorPixelType) colorspace, UNSIGNED_8,
                           (JxlCompressionOption) compression, num_loops, quality, effort, decoding_speed);
    size_t ch = colorspace == 2 ? 4 : colorspace == 1 ? 3 : 1;
    size_t fsz = (size_t) w * h * ch;
    for (int i = 0; i < nframes; i++) {
      std::vector<uint8_t> f(frames + fsz * i, frames + fsz * (i + 1));
      enc.addFrame(f, duration);
    }
    std::vector<uint8_t> dst;
    enc.encode(dst);
    *out = (uint8_t *) malloc(dst.size());
    memcpy(*out, dst.data(), dst.size());
    *out_len = dst.size();
    return 0;
  } catch (std::exception &) {
    return 1;
  }
}

// ---- imagebit / weaver wrappers (unit parity for the pack + rescale kernels) ---------------------------------------
void ref_associate_alpha_rgba8(const uint8_t *src, uint32_t ss, uint8_t *dst, uint32_t ds, uint32_t w, uint32_t h) {
  coder::AssociateAlphaRgba8(src, ss, dst, ds, w, h);
}
void ref_associate_alpha_rgba16(const uint16_t *src, uint32_t ss, uint16_t *dst, uint32_t ds, uint32_t w, uint32_t h, int depth) {
  coder::AssociateAlphaRgba16(src, ss, dst, ds, w, h, depth);
}
void ref_rgba8_to_f16(const uint8_t *src, uint32_t ss, uint16_t *dst, uint32_t ds, uint32_t w, uint32_t h, int attenuate) {
  coder::Rgba8ToF16(src, ss, dst, ds, w, h, attenuate != 0);
}
void ref_rgba16_to_f16(const uint16_t *src, uint32_t ss, uint16_t *dst, uint32_t ds, uint32_t w, uint32_t h, int depth) {
  coder::RgbaU16ToF(src, ss, dst, ds, w, h, depth);
}
void ref_rgba8_to_1010102(const uint8_t *src, uint32_t ss, uint8_t *dst, uint32_t ds, uint32_t w, uint32_t h, int attenuate) {
  coder::Rgba8ToRGBA1010102(src, ss, dst, ds, w, h, attenuate != 0);
}
void ref_rgba16_to_1010102(const uint16_t *src, uint32_t ss, uint8_t *dst, uint32_t ds, uint32_t w, uint32_t h, int depth) {
  coder::Rgba16ToRGBA1010102(src, ss, dst, ds, w, h, depth);
}
void ref_rgba8_to_565(const uint8_t *src, uint32_t ss, uint16_t *dst, uint32_t ds, uint32_t w, uint32_t h, int attenuate) {
  coder::Rgba8To565(src, ss, dst, ds, w, h, attenuate != 0);
}
void ref_rgba16_to_565(const uint16_t *src, uint32_t ss, uint16_t *dst, uint32_t ds, uint32_t w, uint32_t h, int depth) {
  coder::Rgba16To565(src, ss, dst, ds, w, h, depth);
}
void ref_rgba16_to_rgba8(const uint16_t *src, uint32_t ss, uint8_t *dst, uint32_t ds, uint32_t w, uint32_t h, int depth) {
  coder::Rgba16ToRgba8(src, ss, dst, ds, w, h, depth);
}

// returns 0 on success; *out is malloc'ed (ref_free), dims in ow/oh/ostride (bytes)
int ref_weave_u8(const uint8_t *src, uint32_t stride, uint32_t w, uint32_t h, int nw, int nh, int fn, int premul,
                 int mode, uint8_t **out, uint32_t *ow, uint32_t *oh, uint32_t *ostride) {
  ScalingResultU8 r = weave_scale_u8(src, stride, w, h, nw, nh, (ScalingFunction) fn, premul != 0, (WeaveScaleMode) mode);
  if (!r.data) return 1;
  *out = (uint8_t *) malloc(r.length);
  memcpy(*out, r.data, r.length);
  *ow = (uint32_t) r.width;
  *oh = (uint32_t) r.height;
  *ostride = (uint32_t) r.stride;
  weave_scaling_result_free(r);
  return 0;
}
int ref_weave_u16(const uint16_t *src, uint32_t stride_bytes, uint32_t w, uint32_t h, int nw, int nh, int depth,
                  int fn, int premul, int mode, uint16_t **out, uint32_t *ow, uint32_t *oh, uint32_t *ostride_elems) {
  ScalingResultU16 r = weave_scale_u16(src, stride_bytes, w, h, nw, nh, (uintptr_t) depth, (ScalingFunction) fn, premul != 0, (WeaveScaleMode) mode);
  if (!r.data) return 1;
  *out = (uint16_t *) malloc(r.length * 2);
  memcpy(*out, r.data, r.length * 2);
  *ow = (uint32_t) r.width;
  *oh = (uint32_t) r.height;
  *ostride_elems = (uint32_t) r.stride;
  weave_scaling_result16_free(r);
  return 0;
}

uint32_t ref_libjxl_version() { return JxlDecoderVersion(); }

}  // extern "C"

// ---- libjxl-internal inverse transform, called by address ------------------------------------------------------------
// The shipped x86_64 libjxl.so (sha256 25bd94ff…4abe) contains, at file/virtual offset 0x4ff30, the SSE2 instance of
// libjxl's TransformToPixels(AcStrategyType, float* coefficients, float* pixels, size_t pixels_stride, float* scratch)
// (a 27-way jump table on the strategy id).  It is not exported; we reach it relative to the exported
// JxlDecoderVersion (st_value 0x1eb6f0) and refuse to call it unless the prologue bytes match.  This gives the tests
// the reference's exact inverse transform for every block strategy (IDENTITY, DCT2X2, DCT4X4, AFV0-3, DCT4X8, all
// DCT sizes), which the public API only exposes through final pixels.
extern "C" int ref_transform_to_pixels(int strategy, const float *coeffs, size_t ncoef, float *pixels, size_t stride) {
  static const unsigned char kPrologue[20] = {0x55, 0x48, 0x89, 0xe5, 0x41, 0x57, 0x41, 0x56, 0x41, 0x55,
                                             0x41, 0x54, 0x53, 0x48, 0x81, 0xec, 0x78, 0x03, 0x00, 0x00};
  const unsigned char *base = reinterpret_cast<const unsigned char *>(&JxlDecoderVersion) - 0x1eb6f0;
  const unsigned char *fn = base + 0x4ff30;
  if (memcmp(fn, kPrologue, sizeof kPrologue) != 0) return 1;
  if (strategy < 0 || strategy > 26) return 2;
  typedef void (*Fn)(int, float *, float *, size_t, float *);
  size_t n = ncoef < 64 ? 64 : ncoef;
  float *c = (float *) aligned_alloc(64, n * sizeof(float));
  float *scratch = (float *) aligned_alloc(64, (5 * n + 1024) * sizeof(float));
  memcpy(c, coeffs, ncoef * sizeof(float));
  memset(scratch, 0, (5 * n + 1024) * sizeof(float));
  reinterpret_cast<Fn>(const_cast<unsigned char *>(fn))(strategy, c, pixels, stride, scratch);
  free(c);
  free(scratch);
  return 0;
}
