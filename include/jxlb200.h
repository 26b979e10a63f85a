/*
 * jxlb200 — B200-native drop-in for the JPEG XL *decode* path of awxkee/jxl-coder.
 *
 * C ABI of libjxlb200.so.  Every entry point replaces one JNI export of the reference's libjxlcoder.so
 * (jxlcoder/src/main/cpp), minus JNIEnv* / jobject: plain pointers and sizes in, a jxlb_image out.  The JNI (or
 * ctypes / cgo) trampoline a maintainer adds on top is shown in INTEGRATION.md.
 *
 *   reference JNI export (file:line)                                              replacement
 *   Java_..._JxlCoder_decodeSampledImpl           (JniDecoding.cpp:335)           jxlb_decode_sampled
 *   Java_..._JxlCoder_decodeByteBufferSampledImpl (JniDecoding.cpp:363)           jxlb_decode_sampled
 *   Java_..._JxlCoder_getSizeImpl                 (JniDecoding.cpp:396)           jxlb_get_size
 *   Java_..._JxlAnimatedImage_createCoordinator[ByteArray]
 *                                  (JxlAnimatedDecoderCoordinator.cpp:47,95)      jxlb_anim_open
 *   Java_..._JxlAnimatedImage_getNumberOfFrames / getFrameDurationImpl / getLoopsCount / getWidthImpl / getHeightImpl
 *                                  (JxlAnimatedDecoderCoordinator.cpp:139-159,414-425)  jxlb_anim_num_frames / ...
 *   Java_..._JxlAnimatedImage_getFrameImpl        (JxlAnimatedDecoderCoordinator.cpp:162)   jxlb_anim_get_frame
 *   Java_..._JxlAnimatedImage_closeAndReleaseAnimatedImage (…Coordinator.cpp:131)  jxlb_anim_close
 *   (new)                                                                          jxlb_decode_batch
 *
 * Pixels are computed on the GPU only.  If no CUDA device is usable every decode entry point fails with
 * JXLB_ERROR_NO_DEVICE; there is no CPU fallback.
 *
 * Host dependence of lossy output.  The reference's libjxl is a JXL_HIGH_PRECISION=0 SSE2 build whose edge-preserving
 * filter and quant-bias adjustment use the x86 RCPPS instruction, an 11-bit table lookup whose values belong to the CPU
 * vendor.  To return the pictures the reference returns ON THE SAME MACHINE, this library fills its 2048-entry
 * reciprocal table by executing RCPPS on the host CPU when the library is loaded (csrc/numeric_tables.cc: FillRcp11); the
 * GPU then looks reciprocals up in that table.  Consequences: (1) lossy output can differ by 1 LSB on isolated samples
 * between an Intel and an AMD host, exactly as the reference's own output does; (2) on a non-x86 host, or with the
 * environment variable JXLB_EXACT_RCP=1 (read once, at first use), the table holds correctly rounded reciprocals instead
 * -- vendor-independent output that stays within 1 LSB of the reference.  Lossless output never depends on the host.
 */
#ifndef JXLB200_H_
#define JXLB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JXLB_API __attribute__((visibility("default")))

/* Status codes; they mirror the Java exception classes of the reference (JniExceptions.cpp:32-72). */
typedef enum {
  JXLB_OK = 0,
  JXLB_INVALID_JXL = 1,       /* InvalidJXLException: decode returned false (JniDecoding.cpp:70-80) */
  JXLB_INVALID_SIZE = 2,      /* InvalidImageSizeException: >= INT32_MAX bytes (interop/JxlDecoding.cpp:103-109) */
  JXLB_OOM = 3,               /* "Not enough memory to decode this image" (JniDecoding.cpp:81-84) */
  JXLB_BAD_ARG = 4,           /* java.lang.Exception from checkDecodePreconditions (Support.cpp:35-92) */
  JXLB_ERROR = 5,             /* java.lang.Exception("Error: ...") (JniDecoding.cpp:85-89) */
  JXLB_UNSUPPORTED = 6,       /* valid JPEG XL using a coding tool this build does not decode yet */
  JXLB_ERROR_NO_DEVICE = 7,   /* no usable CUDA device / CUDA failure */
  JXLB_NOT_JXL = 8            /* getSize only: not a JPEG XL signature (the reference returns null) */
} jxlb_status;

/* PreferredColorConfig.kt / Support.h:37-44 */
typedef enum {
  JXLB_CONFIG_DEFAULT = 1,
  JXLB_CONFIG_RGBA_8888 = 2,
  JXLB_CONFIG_RGBA_F16 = 3,
  JXLB_CONFIG_RGB_565 = 4,
  JXLB_CONFIG_RGBA_1010102 = 5,
  JXLB_CONFIG_HARDWARE = 6     /* Android-only; always JXLB_ERROR here, as on a device without AHardwareBuffer */
} jxlb_color_config;

/* ScaleMode.kt / SizeScaler.h:36-40 */
typedef enum { JXLB_SCALE_FIT = 1, JXLB_SCALE_FILL = 2, JXLB_SCALE_RESIZE = 3 } jxlb_scale_mode;

/* JxlResizeFilter.kt / XScaler.h:34-45 */
typedef enum {
  JXLB_FILTER_BILINEAR = 1, JXLB_FILTER_NEAREST = 2, JXLB_FILTER_CUBIC = 3, JXLB_FILTER_MITCHELL = 4,
  JXLB_FILTER_LANCZOS = 5, JXLB_FILTER_CATMULL_ROM = 6, JXLB_FILTER_HERMITE = 7, JXLB_FILTER_BSPLINE = 8,
  JXLB_FILTER_HANN = 9, JXLB_FILTER_BICUBIC = 10
} jxlb_resize_filter;

/* Pixel layout of a result, i.e. the Bitmap.Config the reference would create. */
typedef enum { JXLB_FORMAT_RGBA_8888 = 0, JXLB_FORMAT_RGBA_F16 = 1, JXLB_FORMAT_RGB_565 = 2, JXLB_FORMAT_RGBA_1010102 = 3 } jxlb_format;

/* NativeColorSpace.h:10-18 — the ColorSpace.Named the reference tags the Bitmap with when api_level >= 34. */
typedef enum {
  JXLB_CS_NONE = 0, JXLB_CS_SRGB = 1, JXLB_CS_BT2020_PQ = 2, JXLB_CS_BT2020_HLG = 3, JXLB_CS_DISPLAY_P3 = 4,
  JXLB_CS_LINEAR_SRGB = 5, JXLB_CS_DCI_P3 = 6, JXLB_CS_BT709 = 7
} jxlb_color_space;

/* What android.graphics.Bitmap carries.  Callee allocates `data`; release with jxlb_image_free. */
typedef struct {
  void* data;               /* pixels: pinned host memory (device == -1) or device memory of CUDA ordinal `device` */
  uint32_t width, height;
  uint32_t stride_bytes;
  int32_t format;           /* jxlb_format */
  int32_t color_space;      /* jxlb_color_space */
  int32_t premultiplied;    /* 1: colour samples are premultiplied by alpha (ReformatBitmap.cpp:65-77) */
  int32_t device;           /* -1 = host */
  char message[128];        /* human-readable reason when the status is not JXLB_OK */
} jxlb_image;

/* One request of a batch: semantically one decodeSampled call. */
typedef struct {
  const uint8_t* data;
  size_t len;
  int32_t width, height;    /* <= 0 in both: no rescale (JxlCoder.kt:55-62) */
  int32_t color_config;     /* jxlb_color_config */
  int32_t scale_mode;       /* jxlb_scale_mode */
  int32_t filter;           /* jxlb_resize_filter */
} jxlb_request;

typedef struct {
  int32_t api_level;        /* android API level the reference would run on; 0 = 34 (no colour-matrix pass, tag returned) */
  int32_t output_device;    /* -1: results in pinned host memory; >= 0: leave results in device memory of that ordinal */
  int32_t device;           /* CUDA ordinal to decode on; -1 = current device */
  int32_t reserved;
} jxlb_batch_opts;

/* JxlCoder.decodeSampled (byte[] / direct ByteBuffer): the input is copied on entry and never retained. */
JXLB_API int jxlb_decode_sampled(const uint8_t* data, size_t len, int32_t width, int32_t height, int32_t color_config,
                                 int32_t scale_mode, int32_t filter, int32_t api_level, jxlb_image* out);

/* n independent decodeSampled calls decoded together on one GPU; outs[i].message / return code per image in status[i]
   (may be NULL).  Returns JXLB_OK when every image decoded. */
JXLB_API int jxlb_decode_batch(const jxlb_request* reqs, size_t n, jxlb_image* outs, int32_t* status, const jxlb_batch_opts* opts);

/* The same decode, asynchronously.  jxlb_decode_batch_submit parses the requests on the calling thread (the input buffers
   are copied, as the JNI entry copies its byte array on entry -- JniDecoding.cpp:342-345 -- and may be released when it
   returns), starts upload / kernels / downloads on a worker thread and returns a handle (NULL on bad arguments or out of
   memory).  jxlb_decode_batch_collect waits for the batch, fills outs / status exactly like jxlb_decode_batch and frees the
   handle.  A single-threaded caller keeps the GPU and the PCIe link busy by holding two or three batches in flight:
   one synchronous call is a latency chain (its LF stage alone is ~86 ms of serial entropy chains, whatever the batch size).
   Handles may be collected in any order and from any thread; every handle must be collected exactly once. */
typedef struct jxlb_pending jxlb_pending;
JXLB_API jxlb_pending* jxlb_decode_batch_submit(const jxlb_request* reqs, size_t n, const jxlb_batch_opts* opts);
JXLB_API int jxlb_decode_batch_collect(jxlb_pending* pending, jxlb_image* outs, int32_t* status);

/* JxlCoder.getSize: JXLB_OK and (*width, *height), or JXLB_NOT_JXL / JXLB_INVALID_JXL ("null" in the reference). */
JXLB_API int jxlb_get_size(const uint8_t* data, size_t len, uint32_t* width, uint32_t* height);

JXLB_API void jxlb_image_free(jxlb_image* img);

/* TEST HOOK (no counterpart in the reference; used by tests/test_gpu_recon_block.py).  Runs the reconstruction kernels
   (dequantisation, LLF from LF, inverse transform) on ONE synthetic block of `strategy` (0 .. 26, JPEG XL's AcStrategy
   numbering) on CUDA device `device` (-1 = current).  q: [3][8 cy][8 cx] quantised coefficients, X / Y / B, row = vertical
   frequency; lf: [3][cy][cx] dequantised LF samples; out: [3][8 cy][8 cx] XYB samples.  libjxl's encoder never emits
   DCT128 / DCT256 blocks, so no file reaches those transforms: this is how the tests drive them on the GPU. */
JXLB_API int jxlb_test_recon_block(int device, uint32_t strategy, const int16_t* q, const float* lf, uint32_t hf_mul,
                                   uint32_t global_scale, float* out);

/* Prepared batches (throughput interface).  jxlb_batch_prepare parses the requests on the CPU and uploads the
   codestreams + tables, so the inputs are resident in HBM; jxlb_batch_run executes every kernel (entropy decode ->
   packed pixels) and leaves the results in HBM; it can be called repeatedly.  jxlb_batch_fetch copies one result to
   pinned host memory.  status (may be NULL) receives one jxlb_status per request. */
typedef struct jxlb_batch jxlb_batch;
JXLB_API jxlb_batch* jxlb_batch_prepare(const jxlb_request* reqs, size_t n, const jxlb_batch_opts* opts, int32_t* status);
JXLB_API int jxlb_batch_run(jxlb_batch* b);
/* Asynchronous form: jxlb_batch_run_async enqueues one run on the batch's own CUDA stream and returns; jxlb_batch_wait
   waits for every run enqueued so far and returns the status.  Runs of DIFFERENT prepared batches overlap on the GPU
   (the latency-bound LF-group stage of one hides under the throughput kernels of the other); runs of the same batch
   execute in order. */
JXLB_API int jxlb_batch_run_async(jxlb_batch* b);
JXLB_API int jxlb_batch_wait(jxlb_batch* b);
JXLB_API int jxlb_batch_fetch(jxlb_batch* b, size_t index, jxlb_image* out);
/* Device pointer + size of result `index` after jxlb_batch_run (valid until the next run / free). */
JXLB_API const void* jxlb_batch_device_pixels(const jxlb_batch* b, size_t index, size_t* bytes);
/* Device time (ms, CUDA events on the decode stream) of the last run: [0] upload, [1] LF sections, [2] group sections,
   [3] reconstruction phase (per image: LF dequant + smoothing, inverse transforms, filters + colour + pack, interleaved),
   [4] dequant + inverse-transform kernels, [5] filter + colour + pack kernels ([4], [5]: mean of the sampled launches x images),
   [6] download, [7] all kernels. */
JXLB_API void jxlb_batch_stage_ms(const jxlb_batch* b, float* ms8);
/* Same layout, averaged over every run of the batch waited for so far; *runs (may be NULL) = how many. */
JXLB_API void jxlb_batch_stage_ms_mean(const jxlb_batch* b, float* ms8, int32_t* runs);
/* Measurement helpers: jxlb_batch_reset_stats clears the stage-time averages and re-arms the span start;
   jxlb_batch_span_ms = device time (CUDA events) from the start of `first`'s first run since its reset to the end of
   `last`'s latest run (wait for both first) -- the timed region of a run that overlaps several prepared batches. */
JXLB_API void jxlb_batch_reset_stats(jxlb_batch* b);
JXLB_API float jxlb_batch_span_ms(const jxlb_batch* first, const jxlb_batch* last);
JXLB_API void jxlb_batch_free(jxlb_batch* b);

/* JxlAnimatedImage */
typedef struct jxlb_anim jxlb_anim;
JXLB_API jxlb_anim* jxlb_anim_open(const uint8_t* data, size_t len, int32_t color_config, int32_t scale_mode, int32_t filter,
                                   int32_t api_level, int32_t* status);
JXLB_API int32_t jxlb_anim_num_frames(const jxlb_anim* a);
JXLB_API int32_t jxlb_anim_frame_duration_ms(const jxlb_anim* a, int32_t frame);
JXLB_API int32_t jxlb_anim_loops(const jxlb_anim* a);
JXLB_API int32_t jxlb_anim_width(const jxlb_anim* a);
JXLB_API int32_t jxlb_anim_height(const jxlb_anim* a);
JXLB_API int jxlb_anim_get_frame(jxlb_anim* a, int32_t frame, int32_t width, int32_t height, jxlb_image* out);
JXLB_API void jxlb_anim_close(jxlb_anim* a);

/* Introspection for tests / bench: kernels launched by this library so far; per-stage device time of the last batch
   (ms, CUDA events on the decode stream): [0] upload, [1] entropy (LF + groups), [2] reconstruction (LF final + IDCT),
   [3] filters + colour + pack, [4] download, [5] whole batch. */
JXLB_API uint64_t jxlb_kernel_launches(void);
JXLB_API void jxlb_last_batch_timings(float* ms6);
JXLB_API const char* jxlb_version(void);

#ifdef __cplusplus
}
#endif
#endif  /* JXLB200_H_ */
