// TEST INFRASTRUCTURE (oracle): stand-in for the NDK <android/hardware_buffer.h> (types only;
// the HARDWARE colour config is an Android-only output target and stays unavailable here).
#pragma once
#include <cstdint>
typedef struct AHardwareBuffer AHardwareBuffer;
typedef struct { int32_t left, top, right, bottom; } ARect;
typedef struct AHardwareBuffer_Desc {
  uint32_t width, height, layers, format;
  uint64_t usage;
  uint32_t stride, rfu0;
  uint64_t rfu1;
} AHardwareBuffer_Desc;
enum {
  AHARDWAREBUFFER_FORMAT_R8G8B8A8_UNORM = 1,
  AHARDWAREBUFFER_FORMAT_R16G16B16A16_FLOAT = 0x16,
  AHARDWAREBUFFER_USAGE_CPU_READ_OFTEN = 3,
  AHARDWAREBUFFER_USAGE_CPU_WRITE_OFTEN = 3 << 4,
  AHARDWAREBUFFER_USAGE_GPU_SAMPLED_IMAGE = 1 << 8,
  AHARDWAREBUFFER_USAGE_GPU_COLOR_OUTPUT = 1 << 9,
};
