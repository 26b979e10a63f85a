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

// Decodes n requests on CUDA device `device` (-1 = current).  Thread-safe (calls on the same device serialise).
int DecodeBatch(const jxlb_request* reqs, size_t n, int api_level, int device, int output_device, std::vector<DecodedImage>* out,
                BatchTimings* timings);

void FreeImageMemory(void* data, int device);

}  // namespace jxlb
