// Launch wrappers of the CUDA kernels (kernels.cu).  All launches are asynchronous on the given stream.
#pragma once
#include <cuda_runtime.h>
#include "frame.h"
#include "numeric.h"
#include "pack.h"
#include "pixel_stages.h"
#include "vardct_sections.h"

namespace jxlb {

// One serial bitstream section to decode: frame index in the batch + section-specific index.
struct StreamJob {
  uint32_t frame;
  uint32_t index;        // LF group / group index
  uint32_t status_slot;  // entry of FrameDev::status to write
  uint32_t lz_slot;      // 1-based LZ77 window of ScratchLayout::lz_base for this stream's modular data, 0 = none
};

// Per-launch scratch: job j owns [j * bytes_per_job, (j + 1) * bytes_per_job) of `base`, carved into
// arena | weighted-predictor state | non-zero map | HfGlobal arena + permutation scratch (single-section only).
struct ScratchLayout {
  uint8_t* base;
  uint64_t bytes_per_job;
  uint32_t arena_bytes;
  uint32_t wp_ints;
  uint32_t hf_arena_bytes;   // single-section frames only
  uint32_t max_local_nodes;
  // LZ77 windows for the modular streams of frames whose global code uses LZ77 (libjxl's effort-1 lossless encoder):
  // lz_entries (a power of two >= the symbols of one stream, at most 2^20 as in libjxl) uint32 each
  uint8_t* lz_base;
  uint32_t lz_entries;
};

void LaunchSingleSectionFrames(const FrameDev* frames, const StreamJob* jobs, uint32_t njobs, NaturalOrders nat,
                               ScratchLayout scratch, cudaStream_t stream);
// After all entropy kernels of a batch: marks the frames with a failed section (FrameDev::frame_bad).
void LaunchFrameStatus(const FrameDev* frames, uint32_t nframes, cudaStream_t stream);
void LaunchLfGroups(const FrameDev* frames, const StreamJob* jobs, uint32_t njobs, ScratchLayout scratch, cudaStream_t stream);
void LaunchPassGroups(const FrameDev* frames, const StreamJob* jobs, uint32_t njobs, NaturalOrders nat, ScratchLayout scratch,
                      cudaStream_t stream);

// Lane-parallel AC decode (kernels_ac.cu): a CTA of 128 lanes decodes up to 128 groups of ONE frame.
struct AcCtaJob {
  uint32_t frame;
  uint32_t first_group;
  uint32_t ngroups;
  uint32_t pad;
};
// Rescale (kernels_resize.cu): device views of a ResizePlan's axis tables and of the three images involved.
struct ResizeAxisDev {
  const uint32_t* start;
  const uint32_t* count;
  const int16_t* weights;
  uint32_t taps;
};
struct ResizeDev {
  const uint8_t* src;      // RGBA8, src_stride bytes per row
  uint32_t src_stride, src_w, src_h, scaled_w, scaled_h;
  bool has_v, has_h;       // a pass whose size does not change is skipped
  bool nearest;            // one gather pass: v.start / h.start are the source row / column of every output row / column
  bool premultiply;        // colour * alpha (rounded / 255) on the way in, divided back on the way out
  ResizeAxisDev v, h;
  uint8_t* mid;            // [scaled_h][src_w] RGBA8
  uint8_t* scaled;         // [scaled_h][scaled_w] RGBA8
};
// Returns the pointer holding the result: r.scaled, r.mid (horizontal pass skipped; row stride src_w * 4) or r.src.
const uint8_t* LaunchResize(const ResizeDev& r, cudaStream_t stream);

// Inverse squeeze of a frame's lossy extra channels (squeeze.h) into f.mod, after every group section has been decoded;
// steps_host: the host copy of f.sq_steps (launch geometry).
void LaunchUnsqueeze(const FrameDev& f, const SqStep* steps_host, cudaStream_t stream);

// One step of frame composition (kernels_post.cu): the frame `fg` (fw x fh straight RGBA8 / RGBA16 at (x0, y0), may stick out
// of the canvas) blended over the canvas `bg` (null = empty canvas) into `out` (and `save`: the reference slot the frame
// is stored in, may be null).  Blend modes as coded: 0 kReplace, 1 kAdd, 2 kBlend, 3 kAlphaWeightedAdd, 4 kMul.
struct CompositeParams {
  const float4* bg;      // reference slot the frame is blended over: cw x ch float RGBA in [0, 1], null = empty canvas
  const uint8_t* fg;     // the frame: straight RGBA8 / RGBA16
  uint8_t* out;          // the picture as integers (last step only; null for intermediate frames)
  float4* save;          // reference slot the result is stored in (null = not stored)
  uint32_t fg_stride, fw, fh;
  int32_t x0, y0;
  uint32_t cw, ch, canvas_stride;
  uint32_t bits16, has_alpha, alpha_premultiplied, mode_color, mode_alpha, clamp;
  const float* dither;   // the 32x32 dither table for the LAST step of an 8-bit picture, else null
  uint32_t orientation;  // codestream orientation (the dither pattern is indexed by the flipped position, pixel_stages.h)
};
void LaunchComposite(const CompositeParams& p, cudaStream_t stream);

// Lays an fw x fh frame at (x0, y0) over a cleared cw x ch canvas (colour 0, alpha fill_alpha in the sample depth).
void LaunchPlace(const uint8_t* src, uint32_t src_stride, uint32_t fw, uint32_t fh, uint32_t bpp, int32_t x0, int32_t y0, uint32_t fill_alpha,
                 uint8_t* dst, uint32_t dst_stride, uint32_t cw, uint32_t ch, cudaStream_t stream);

// Codestream orientation (EXIF numbering 2..8) applied to an interleaved image of bpp (4 or 8) bytes per pixel; src is w x h,
// dst is w x h (2..4) or h x w (5..8).
void LaunchOrient(const uint8_t* src, uint32_t src_stride, uint32_t w, uint32_t h, uint32_t bpp, uint32_t orientation, uint8_t* dst,
                  uint32_t dst_stride, cudaStream_t stream);

// api_level < 34 colour pass (kernels_post.cu), in place on straight RGBA8 (or RGBA16: bits16); plan_dev: a ColorMatrixPlan
// in device memory, followed at the next 256-byte boundary by a ColorMatrixTables16 when bits16.  tonemap: PQ / HLG source
// (Rec.2408 tone mapping); row_first: `height` words of device scratch for it.
struct ColorMatrixPlan;
void LaunchColorMatrix(uint8_t* img, uint32_t stride, uint32_t width, uint32_t height, const ColorMatrixPlan* plan_dev, bool tonemap, bool bits16,
                       uint32_t* row_first, cudaStream_t stream);

// Sets `bytes` (a multiple of 16, 16-byte aligned) to the repeated 32-bit value with a kernel.
void LaunchFill(void* p, size_t bytes, uint32_t value32, cudaStream_t stream);
void LaunchBuildGroupBlocks(const FrameDev* frames, const StreamJob* jobs, uint32_t njobs, cudaStream_t stream);
uint32_t AcLaneSmemBytes(uint32_t code_bytes);
uint32_t AcGroupsPerCta();  // groups one CTA of the lane-parallel AC kernel decodes (128 or 256)
void LaunchAcLanes(const FrameDev* frames, const AcCtaJob* jobs, uint32_t njobs, NaturalOrders nat, uint32_t smem_code_bytes,
                   bool fast, cudaStream_t stream);
void LaunchGroupModular(const FrameDev* frames, const StreamJob* jobs, uint32_t njobs, ScratchLayout scratch, cudaStream_t stream);

// Numeric stages of one VarDCT frame (FrameDev passed by value).
void LaunchLfFinal(const FrameDev& f, cudaStream_t stream);
void LaunchRecon(const FrameDev& f, const NumericTables* nt_dev, cudaStream_t stream);
// Contained blocks of every 64x64 region through the persistent TMA kernel (kernels_recon.cu); false = not usable here.
bool LaunchReconTma(const FrameDev& f, const NumericTables* nt_dev, cudaStream_t stream);
// Gaborish + EPF stages as configured; returns which buffer (0 = xyb0, 1 = xyb1) holds the result.
int LaunchFilters(const FrameDev& f, const NumericTables* nt_dev, cudaStream_t stream);
void LaunchColor(const FrameDev& f, const ColorParams& cp, const NumericTables* nt_dev, const float* src, OutputDesc out,
                 cudaStream_t stream);
void LaunchModularToRgba(const FrameDev& f, OutputDesc out, cudaStream_t stream);
void LaunchModularGlobalInverse(const FrameDev& f, cudaStream_t stream);
void LaunchScatterGlobalPlanes(const FrameDev& f, cudaStream_t stream);
// 2x upsampling of the filtered XYB planes of a frame coded at half resolution: src (f geometry) -> dst [3][up_h][up_stride]
void LaunchUpsample2(const FrameDev& f, const float* src, float* dst, uint32_t up_stride, uint32_t up_h, cudaStream_t stream);
// ... and of its alpha plane (coded int32 samples of `bits` bits -> floats in [0, 1], [up_h][up_stride]; OutputDesc::alpha_float)
void LaunchUpsampleAlpha2(const FrameDev& f, const int32_t* src, uint32_t bits, int32_t* dst, uint32_t up_stride,
                          cudaStream_t stream);
// Fused Gaborish + EPF + colour + pack (kernels_filter.cu): XYB planes in f.xyb0 -> packed pixels.
void LaunchFilterColorPack(const FrameDev& f, const ColorParams& cp, const NumericTables* nt_dev, const OutputDesc& od,
                           const PackParams& pack, cudaStream_t stream);
// Alpha association + packing to the requested Bitmap format (pack.h).
void LaunchPack(const PackParams& p, cudaStream_t stream);

// Number of kernel launches issued by this module since process start (bench.py reports it as gpu_launches).
uint64_t KernelLaunchCount();

}  // namespace jxlb
