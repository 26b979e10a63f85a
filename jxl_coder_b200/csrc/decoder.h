// Batch decoder: host-side orchestration of one GPU (parse -> plan -> upload -> kernels -> download).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/jxlb200.h"

namespace jxlb {

struct DecodedImage {
  int status = JXLB_OK;          // jxlb_status
  std::string message;
  void* data = nullptr;          // pinned host or device memory (owned by the caller after a successful decode)
  uint32_t width = 0, height = 0, stride_bytes = 0;
  int format = JXLB_FORMAT_RGBA_8888;
  int color_space = JXLB_CS_NONE;
  int premultiplied = 0;
  int device = -1;
};

struct BatchTimings {
  float ms[6] = {0, 0, 0, 0, 0, 0};
};

// Decodes n requests on CUDA device `device` (-1 = current).  Thread-safe; concurrent calls on the same device take
// different decode slots (streams + buffers) and overlap on the GPU.
// frame_index (may be null): per request, which DISPLAYED frame (frame types regular / skip-progressive, in codestream
// order) to decode; -1 or null = the last frame, as the still-image entry points do.  Frames that depend on earlier ones
// (blend modes, crops over a kept canvas, reference slots) are composed on the GPU from the chain of frames they need;
// only references stored before the colour transform (patches) report JXLB_UNSUPPORTED.
int DecodeBatch(const jxlb_request* reqs, size_t n, int api_level, int device, int output_device, std::vector<DecodedImage>* out,
                BatchTimings* timings, const int32_t* frame_index = nullptr);

void FreeImageMemory(void* data, int device);

// Asynchronous form of DecodeBatch: SubmitBatch parses the requests (the input buffers may be released when it returns),
// starts the device half on a worker thread and returns a handle; CollectBatch waits for it, hands over the results and
// frees the handle.  Several submitted batches overlap on the GPU like concurrent DecodeBatch calls do.
struct PendingBatch;
PendingBatch* SubmitBatch(const jxlb_request* reqs, size_t n, int api_level, int device, int output_device,
                          const int32_t* frame_index = nullptr);
int CollectBatch(PendingBatch* p, std::vector<DecodedImage>* out, BatchTimings* timings);

}  // namespace jxlb

// ---- prepared batches: parse + upload once ("inputs resident in HBM"), then run the kernels any number of times with
// the results left in HBM.  Used by the benchmark's device-resident measurement and by callers that keep pixels on the GPU.
#include <cuda_runtime.h>
namespace jxlb {
struct Batch;
Batch* PrepareBatch(const jxlb_request* reqs, size_t n, int api_level, int device, std::vector<int>* status);
int RunBatch(Batch* b, bool sync);
// Waits for the asynchronous runs issued with RunBatch(b, false) and resolves statuses.
int WaitBatch(Batch* b);
// Mean stage times over every run collected so far (same layout as BatchStageMs).
void BatchStageMsMean(const Batch* b, float* ms8, int* runs);
void ResetBatchStats(Batch* b);
float BatchSpanMs(const Batch* first, const Batch* last);
int FetchBatchImage(Batch* b, size_t i, DecodedImage* out);
// ms8: [0] upload, [1] LF sections, [2] group sections, [3] reconstruction phase, [4] dequant+inverse transform kernels (sampled),
//      [5] filters+colour+pack, [6] download, [7] all kernels
void BatchStageMs(const Batch* b, float* ms8);
const void* BatchDevicePixels(const Batch* b, size_t i, size_t* bytes);
cudaStream_t BatchStream(const Batch* b);
void FreeBatch(Batch* b);
// Test hook (include/jxlb200.h: jxlb_test_recon_block).
int TestReconBlock(int device, uint32_t strategy, const int16_t* q, const float* lf, uint32_t hf_mul, uint32_t global_scale, float* out);
}  // namespace jxlb
