// Memory plan of one frame: a "const" region (codestream + host-parsed tables, uploaded once) and a "work" region
// (planes the kernels produce).  The same plan is used with HBM pointers by the decoder and with host pointers by the
// CPU-side unit tests of the shared section decoders (tests/hostemu).
#pragma once
#include <cstdint>
#include <vector>
#include "frame_parser.h"

namespace jxlb {

struct FramePlan {
  // const region
  size_t const_bytes = 0;
  size_t off_cs = 0, off_sec_begin = 0, off_sec_end = 0, off_bctx_map = 0, off_tree = 0, off_tree_code = 0, off_ac_code = 0,
         off_order_pool = 0, off_blockinfo_off = 0, off_sq_ch = 0, off_sq_steps = 0, off_sq_global = 0, off_meta = 0, off_pass_table = 0, off_global_planes = 0;
  // work region
  size_t work_bytes = 0;
  size_t off_lf_quant = 0, off_xfromy = 0, off_bfromy = 0, off_sharp_i32 = 0, off_blockinfo = 0, off_nb_blocks = 0,
         off_extra_prec = 0, off_cell_strategy = 0, off_cell_hfmul = 0, off_cell_sharp = 0, off_cell_off = 0, off_group_blocks = 0, off_group_nblocks = 0, off_group_ac_end = 0, off_coef = 0, off_lf = 0, off_large_list = 0,
         off_xyb0 = 0, off_xyb1 = 0, off_mod = 0, off_status = 0, off_sq_buf = 0, off_frame_bad = 0;
  size_t coef_bytes = 0;   // zero-filled before the AC decode
  size_t up_bytes = 0;     // upsampled XYB planes of a frame coded at half resolution (same ring slot, after two plane sets)
  size_t xyb_bytes = 0;    // one XYB f32 plane set [3][plane_h][plane_stride]; NOT part of the work region: the decoder
                           // binds FrameDev::xyb0 (and xyb1 for the unfused debug path) to a small ring of such buffers
  uint32_t num_streams = 0;  // status entries: [lf groups][groups] (+1 for the single-section chain)
  FrameDev proto{};          // all scalar fields filled; pointers null
};

// Computes the plan.  cs_padded_bytes = size of the padded codestream buffer.
// True when JXLB_SIMPLE_FILTERS is set: restoration filters run as separate per-stage kernels (debug aid) instead of the
// fused filter + colour + pack kernel.
bool UseUnfusedFilters();
void MakeFramePlan(const ImageMetadata& md, const FrameHeader& fh, const FrameGlobals& g, size_t cs_padded_bytes, FramePlan* plan);

// Serialises the const region into `dst` (plan.const_bytes bytes).
void FillConstRegion(const FramePlan& plan, const uint8_t* cs_padded, const FrameHeader& fh, const FrameGlobals& g, uint8_t* dst);

// Produces the FrameDev with pointers into the given regions (host or device addresses).
FrameDev BindFrameDev(const FramePlan& plan, const uint8_t* const_base, uint8_t* work_base);

}  // namespace jxlb
