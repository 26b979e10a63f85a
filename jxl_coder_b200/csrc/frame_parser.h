// Host-side parsing of everything that is global to an image or a frame: container, SizeHeader, ImageMetadata,
// FrameHeader, TOC, LfGlobal and HfGlobal.  These sections are tiny, serial and bit-twiddly; the per-group sections
// they describe are decoded on the GPU.  Mirrors what the reference gets from libjxl's JXL_DEC_BASIC_INFO /
// JXL_DEC_COLOR_ENCODING events (/root/reference/jxlcoder/src/main/cpp/interop/JxlDecoding.cpp:81-144) and from
// JxlDecoderProcessInput's header handling.  Format digest: SURVEY.md App. B.1-B.5, B.7.
#pragma once
#include "recycle_alloc.h"
#include <cstdint>
#include <string>
#include <vector>
#include "frame.h"
#include "squeeze.h"

namespace jxlb {

// C-ABI level error classes (include/jxlb200.h mirrors these values).
enum ParseStatus : int {
  kParseOk = 0,
  kParseNotJxl = 1,         // signature mismatch
  kParseInvalid = 2,        // corrupt / truncated
  kParseUnsupported = 3,    // valid JPEG XL feature outside this build's coverage
};

struct ExtraChannelInfo {
  uint32_t type = 0;        // 0 = alpha
  uint32_t bits = 8, exp_bits = 0;
  bool is_float = false;
  uint32_t dim_shift = 0;
  bool alpha_premultiplied = false;
};

struct ColorEncoding {
  bool all_default = true;
  bool want_icc = false;
  uint32_t color_space = 0;   // 0 RGB, 1 grey, 2 XYB, 3 unknown       (jxl/color_encoding.h enums)
  uint32_t white_point = 1;   // 1 D65
  uint32_t primaries = 1;     // 1 sRGB, 9 2100, 11 P3
  bool have_gamma = false;
  uint32_t gamma_u24 = 0;
  uint32_t transfer = 13;     // 13 sRGB, 8 linear, 16 PQ, 18 HLG, 1 709, 17 DCI
  uint32_t rendering_intent = 1;
  int32_t white_xy[2] = {0, 0};
  int32_t prim_xy[3][2] = {{0, 0}, {0, 0}, {0, 0}};
};

struct ImageMetadata {
  uint32_t xsize = 0, ysize = 0;
  uint32_t orientation = 1;
  bool have_intrinsic_size = false, have_preview = false, have_animation = false;
  uint32_t tps_num = 0, tps_den = 0, num_loops = 0;
  bool have_timecodes = false;
  uint32_t bits_per_sample = 8, exp_bits = 0;
  bool float_samples = false;
  bool modular_16bit = true;
  std::vector<ExtraChannelInfo> extra;
  bool xyb_encoded = true;
  ColorEncoding color;
  float intensity_target = 255.f, min_nits = 0.f, linear_below = 0.f;
  bool relative_to_max_display = false;
  bool default_transform = true;
  float opsin_inverse[9];
  float opsin_bias[3];
  float quant_bias[3];
  float quant_bias_numerator;
  bool custom_upsampling = false;
  int alpha_channel() const {
    for (size_t i = 0; i < extra.size(); ++i)
      if (extra[i].type == 0) return (int) i;
    return -1;
  }
};

struct BlendingInfo {
  uint32_t mode = 0, alpha_channel = 0, source = 0;
  bool clamp = false;
};

struct FrameHeader {
  uint32_t frame_type = 0;   // 0 regular, 1 LF, 2 reference-only, 3 skip-progressive
  uint32_t encoding = 0;     // 0 VarDCT, 1 modular
  uint64_t flags = 0;
  bool do_ycbcr = false;
  uint32_t jpeg_upsampling[3] = {0, 0, 0};
  uint32_t upsampling = 1;
  std::vector<uint32_t> ec_upsampling;
  uint32_t group_size_shift = 1;
  uint32_t x_qm_scale = 3, b_qm_scale = 2;
  uint32_t num_passes = 1;
  uint32_t pass_shift[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // coefficients of pass p are coded >> pass_shift[p] (0 for the last)
  uint32_t num_ds = 0, pass_downsample[4] = {1, 1, 1, 1}, pass_last[4] = {0, 0, 0, 0};  // resolution a pass completes (modular shifts)
  uint32_t lf_level = 0;
  bool have_crop = false;
  int32_t x0 = 0, y0 = 0;
  uint32_t width = 0, height = 0;   // frame size in image pixels (before upsampling division)
  BlendingInfo blend;
  std::vector<BlendingInfo> ec_blend;
  uint32_t duration = 0, timecode = 0;
  bool is_last = true;
  uint32_t save_as_reference = 0;
  bool save_before_ct = false;
  RestorationFilter rf;
  // derived
  uint32_t coded_w = 0, coded_h = 0, group_dim = 256;
  uint32_t ngx = 0, ngy = 0, nlfx = 0, nlfy = 0, num_groups = 0, num_lf_groups = 0, toc_entries = 0;
  // TOC in logical order, absolute bit positions inside the codestream
  std::vector<uint64_t> sec_bit_begin, sec_bit_end;
  uint64_t end_byte = 0;            // first byte after this frame
};

// Result of parsing LfGlobal / HfGlobal on the host; blobs are copied to HBM verbatim.
struct FrameGlobals {
  float lf_dequant[3] = {1.f / 4096, 1.f / 512, 1.f / 256};
  uint32_t global_scale = 0, quant_lf = 0;
  BlockCtxMap bctx{};
  std::vector<uint8_t> bctx_map;
  CflParams cfl{84, 0.f, 1.f, 128, 128};
  bool has_global_tree = false;
  ByteVec tree_blob;     // TreeNode[num_nodes]
  uint32_t tree_nodes = 0, tree_uses_wp = 0, tree_max_property = 0;
  ByteVec tree_code;     // code blob for the tree's leaf contexts
  uint64_t global_modular_bit = 0;    // where the global modular GroupHeader starts
  ModularHeader global_mh{};          // parsed for multi-section frames with a modular image
  ChannelPlan chplan{};               // ... its channel list (palette meta channels first)
  std::vector<int32_t> meta_data;     // ... and the palette colours, decoded on the host (they live in the global stream)
  // ... and, for a multi-section frame no larger than one group (a small progressive picture), its modular channels
  // themselves, which then live in the global stream too: [chplan.ncoded][coded_h][coded_w], decoded on the host
  std::vector<int32_t> global_planes;
  // squeezed extra channels (squeeze.h): channel pyramid, inverse steps, and the samples of the channels that live in
  // the global stream (decoded on the host: they are at most group_dim x group_dim), concatenated in channel order
  bool squeeze = false;
  SqueezeLayoutOut sq;
  uint32_t sq_global = 0;
  uint64_t sq_end_bit = 0;            // first bit after the global modular stream (single-section frames resume here)
  std::vector<int32_t> sq_global_data;
  // HfGlobal
  bool hf_parsed = false;
  uint32_t num_hf_presets = 1, used_orders = 0;
  U16Vec order_pool;
  OrderTableIndex orders{};
  ByteVec ac_code;
  // progressive frames: the tables of passes 1 .. num_passes - 1 (pass 0 uses the fields above)
  struct ExtraPass {
    uint32_t used_orders = 0;
    OrderTableIndex orders{};
    U16Vec order_pool;
    ByteVec ac_code;
  };
  std::vector<ExtraPass> extra_passes;
  uint64_t hf_global_end_bit = 0;
};

// Extracts the codestream from a bare (FF 0A) or boxed (ISOBMFF) file into `out`, zero-padded by 16 bytes and a
// multiple of 4 long.  *cs_len = unpadded length.
int ExtractCodestream(const uint8_t* data, size_t len, ByteVec* out, size_t* cs_len);

// Parses SizeHeader + ImageMetadata (+ICC skip) and leaves *frame_bit at the first frame header.
int ParseImageHeader(const uint8_t* cs, size_t cs_padded, size_t cs_len, ImageMetadata* md, uint64_t* frame_bit, std::string* err);

// Parses a frame header + TOC starting at bit position frame_bit.
int ParseFrameHeader(const uint8_t* cs, size_t cs_padded, size_t cs_len, const ImageMetadata& md, uint64_t frame_bit,
                     FrameHeader* fh, std::string* err);

// Parses LfGlobal (up to the global modular GroupHeader) and, for multi-section VarDCT frames, HfGlobal.
int ParseFrameGlobals(const uint8_t* cs, size_t cs_padded, const ImageMetadata& md, const FrameHeader& fh, FrameGlobals* g,
                      std::string* err);

// Natural coefficient order of a block covering cx × cy cells (positions index the 8·min × 8·max array).
void NaturalCoeffOrder(uint32_t cx, uint32_t cy, std::vector<uint32_t>* out);

}  // namespace jxlb
