// Common definitions for code shared between the host-side stream parser and the CUDA kernels.
//
// The bit-level parts of the JPEG XL decode path (bit reader, ANS / prefix entropy decoder, MA-tree modular decoder)
// are needed twice: on the host for the small frame-global sections (TOC permutation, LfGlobal, HfGlobal), and on the
// device for everything that scales with pixels (LF groups, AC groups, modular groups).  They are written once as
// JXLB_HD functions over position-independent tables, so the tables the host builds can be copied to HBM verbatim
// and streams with *local* trees/codes can build theirs on the device.
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __CUDACC__
#define JXLB_HD __host__ __device__ __forceinline__
#define JXLB_HD_NOINLINE inline __host__ __device__ __noinline__
#define JXLB_D __device__ __forceinline__
#else
#define JXLB_HD inline
#define JXLB_HD_NOINLINE inline
#define JXLB_D inline
#endif

// Address-space hints for pointers whose provenance the compiler cannot see (they travel through structs and
// non-inlined functions): with them the device code uses LDS / LDG with 32-bit shared addresses instead of generic LD.
#if defined(__CUDA_ARCH__) && !defined(JXLB_NO_HINTS) && !defined(JXLB_NO_HINT_SHARED)
#define JXLB_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
#else
#define JXLB_ASSUME_SHARED(p) ((void) 0)
#endif
#if defined(__CUDA_ARCH__) && !defined(JXLB_NO_HINTS) && !defined(JXLB_NO_HINT_GLOBAL)
#define JXLB_ASSUME_GLOBAL(p) __builtin_assume(__isGlobal(p))
#else
#define JXLB_ASSUME_GLOBAL(p) ((void) 0)
#endif

namespace jxlb {

// Stream-level status codes (device kernels write these per stream; host maps them to the C-ABI error codes).
enum StreamStatus : int32_t {
  kOk = 0,
  kErrTruncated = 1,      // read past the end of the section
  kErrBadStream = 2,      // invalid syntax / failed ANS final-state check
  kErrUnsupported = 3,    // valid JPEG XL this build does not decode yet
  kErrScratch = 4,        // a device scratch arena was too small
};

JXLB_HD int CeilLog2(uint32_t x) {
  int r = 0;
  while ((1u << r) < x && r < 32) ++r;
  return x <= 1 ? 0 : r;
}
JXLB_HD int FloorLog2(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return 31 - __clz(x);
#else
  return 31 - __builtin_clz(x);
#endif
}
JXLB_HD int32_t UnpackSigned(uint32_t v) { return (int32_t) ((v >> 1) ^ (~(v & 1) + 1)); }

}  // namespace jxlb
