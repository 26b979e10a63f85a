// Plain-data description of one JPEG XL frame as the CUDA kernels see it, plus the VarDCT constant tables
// (block strategies, coefficient-order ids, context tables).  Shared by host and device.
// Format digest: SURVEY.md App. B.2-B.7.
#pragma once
#include "hd.h"
#include "squeeze.h"
#include "entropy.h"
#include "modular.h"

namespace jxlb {

static constexpr int kNumStrategies = 27;
static constexpr int kNumOrders = 13;
static constexpr int kNumQuantTables = 17;
static constexpr uint32_t kGroupDim = 256;       // VarDCT group size in pixels
static constexpr uint32_t kGroupCells = 32;      // ... in 8x8 cells
static constexpr uint32_t kLfGroupCells = 256;   // LF group size in cells
static constexpr uint32_t kZeroDensityContexts = 458;
static constexpr uint32_t kNonZeroBuckets = 37;
static constexpr uint32_t kContextsPerBlockCtx = 495;

// cells covered (x, y), coefficient-order id, quant-table id per strategy (App. B.7 strategy table)
// The tables are packed into 64-bit literals (4 or 8 bits per strategy): a local array indexed at run time would be
// built on the stack of every device thread that calls these.
//   cells x: {1, 1, 1, 1, 2, 4, 1, 2, 1, 4, 2, 4, 1, 1, 1, 1, 1, 1, 8, 4, 8, 16, 8, 16, 32, 16, 32}  (stored as log2)
//   cells y: {1, 1, 1, 1, 2, 4, 2, 1, 4, 1, 4, 2, 1, 1, 1, 1, 1, 1, 8, 8, 4, 16, 16, 8, 32, 32, 16}
//   order:   {0, 1, 1, 1, 2, 3, 4, 4, 5, 5, 6, 6, 1, 1, 1, 1, 1, 1, 7, 8, 8, 9, 10, 10, 11, 12, 12}
//   quant:   {0, 1, 2, 3, 4, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 10, 10, 11, 12, 12, 13, 14, 14, 15, 16, 16}
JXLB_HD uint32_t Packed4(uint32_t s, uint64_t lo, uint64_t hi) { return (uint32_t) ((s < 16 ? lo >> (4 * s) : hi >> (4 * (s - 16))) & 15u); }
JXLB_HD uint32_t StrategyCellsXLog2(uint32_t s) { return Packed4(s, 0x212010210000ull, 0x54543432300ull); }
JXLB_HD uint32_t StrategyCellsYLog2(uint32_t s) { return Packed4(s, 0x120201210000ull, 0x45534423300ull); }
JXLB_HD uint32_t StrategyCellsX(uint32_t s) { return 1u << StrategyCellsXLog2(s); }
JXLB_HD uint32_t StrategyCellsY(uint32_t s) { return 1u << StrategyCellsYLog2(s); }
JXLB_HD uint32_t StrategyOrder(uint32_t s) { return Packed4(s, 0x1111665544321110ull, 0xccbaa988711ull); }
JXLB_HD uint32_t StrategyQuantTable(uint32_t s) {
  const uint64_t w = s < 8 ? 0x606050403020100ull : s < 16 ? 0xa0a090908080707ull : s < 24 ? 0xe0e0d0c0c0b0a0aull : 0x10100full;
  return (uint32_t) (w >> (8 * (s & 7u))) & 0xFFu;
}
// a representative strategy for each coefficient-order id
JXLB_HD uint32_t OrderRepresentative(uint32_t o) {
  const uint8_t k[kNumOrders] = {0, 1, 4, 5, 6, 8, 10, 18, 19, 21, 22, 24, 25};
  return k[o];
}
// a representative strategy for each quant table (both orientations share the table)
JXLB_HD uint32_t QuantTableRepresentative(uint32_t q) {
  const uint8_t k[kNumQuantTables] = {0, 1, 2, 3, 4, 5, 6, 8, 10, 12, 14, 18, 19, 21, 22, 24, 25};
  return k[q];
}

// Zero-density context tables (App. B.7: FREQ / NZ)
JXLB_HD uint32_t ZeroDensityFreqCtx(uint32_t k) {
  const uint8_t t[64] = {0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 15, 16, 16, 17, 17, 18, 18, 19, 19, 20, 20, 21, 21, 22, 22,
                         23, 23, 23, 23, 24, 24, 24, 24, 25, 25, 25, 25, 26, 26, 26, 26, 27, 27, 27, 27, 28, 28, 28, 28, 29, 29, 29, 29, 30, 30, 30, 30};
  return t[k];
}
JXLB_HD uint32_t ZeroDensityNnzCtx(uint32_t n) {
  const uint8_t t[64] = {0, 0, 31, 62, 62, 93, 93, 93, 93, 123, 123, 123, 123, 152, 152, 152, 152, 152, 152, 152, 152, 180, 180, 180, 180, 180,
                         180, 180, 180, 180, 180, 180, 180, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206,
                         206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206, 206};
  return t[n];
}

// ---- block-context map (LfGlobal) ----
static constexpr int kMaxThresholds = 15;
struct BlockCtxMap {
  int32_t lf_thr[3][kMaxThresholds];  // X, Y, B
  uint32_t num_lf_thr[3];
  uint32_t qf_thr[kMaxThresholds];
  uint32_t num_qf_thr;
  uint32_t num_lf_ctx;                // product of (num_lf_thr[c] + 1)
  uint32_t num_ctx;                   // nb_block_ctx = max(map) + 1
  uint32_t map_size;
  // ctx_map bytes live in FrameDev::bctx_map (host: std::vector)
};

struct CflParams {
  uint32_t colour_factor;
  float base_x, base_b;
  uint32_t x_factor_lf, b_factor_lf;
};

struct RestorationFilter {
  uint8_t gab, epf_iters, gab_custom, pad;
  float gab_w1[3], gab_w2[3];
  float epf_sharp_lut[8];
  float epf_channel_scale[3];
  float epf_quant_mul, epf_pass0_sigma_scale, epf_pass2_sigma_scale, epf_border_sad_mul, epf_sigma_for_modular;
};

// Coefficient order tables: for order id o and channel c (0=X 1=Y 2=B), order[k] = position of the k-th coded
// coefficient inside the block's 8·min × 8·max coefficient array.  Offsets (in uint16/uint32 entries) per (o, c).
struct OrderTableIndex {
  uint32_t offset[kNumOrders][3];  // into the uint32_t order pool; 0xFFFFFFFF if the order cannot occur
};

// Tables of one further pass of a progressive frame (passes 1 .. num_passes - 1), in the const region; the blobs are
// addressed relative to the start of the PassDev array.
struct PassDev {
  OrderTableIndex orders;
  uint32_t used_orders, shift;
  uint64_t ac_code_rel, order_pool_rel;
};

// Everything global to one frame that the device needs (POD; device pointers filled by the decoder).
struct FrameDev {
  // geometry of the coded frame
  uint32_t width, height;          // pixels
  uint32_t orientation;            // codestream orientation (1..8): the 8-bit dither pattern is indexed by OUTPUT position
  uint32_t dither_x0, dither_y0;   // position of a cropped frame on the canvas, mod 32 (the dither pattern is canvas-anchored)
  uint32_t w8, h8;                 // 8x8 cells
  uint32_t w64, h64;               // CfL tiles
  uint32_t ngx, ngy, nlfx, nlfy;
  uint32_t num_groups, num_lf_groups, num_passes;
  uint32_t group_dim;              // 256 for VarDCT, 128 << shift for modular
  uint32_t encoding;               // 0 VarDCT, 1 modular
  uint32_t flags;
  uint32_t single_section;         // toc_entries == 1
  uint32_t x_qm_scale, b_qm_scale;
  // bitstream (HBM copy of the padded codestream) and the section table in logical order
  const uint8_t* cs;
  uint64_t cs_bytes;               // padded size, multiple of 4
  const uint64_t* sec_bit_begin;   // [toc_entries]
  const uint64_t* sec_bit_end;
  uint32_t toc_entries;
  // LfGlobal
  float lf_dequant[3];             // X, Y, B
  uint32_t global_scale, quant_lf;
  BlockCtxMap bctx;
  const uint8_t* bctx_map;
  CflParams cfl;
  // global MA tree + its code (may be absent)
  const TreeNode* global_tree;
  uint32_t global_tree_nodes, global_tree_uses_wp, global_tree_max_property;
  const uint8_t* global_code;      // code blob
  uint64_t global_modular_bit;     // absolute bit position of the global modular stream's GroupHeader
  // HfGlobal (multi-section frames: parsed on the host)
  uint32_t num_hf_presets;
  uint32_t used_orders;
  const uint8_t* ac_code;          // code blob
  const uint16_t* order_pool;
  OrderTableIndex orders;
  const PassDev* pass_table;       // progressive frames: tables of passes 1 ..; pass 0 uses the fields above
  uint32_t pass_shift0, pass_pad;  // coefficients of pass 0 are coded >> pass_shift0
  uint8_t pass_min_shift[12], pass_max_shift[12];  // modular channels a pass carries: min_shift <= min(hshift, vshift) <= max_shift
  RestorationFilter rf;
  // extra channels / modular image
  uint32_t num_mod_channels;       // channels of the frame's modular image (colour for modular frames + extras)
  uint32_t num_color_mod_channels; // 0 for VarDCT, 1 or 3 for modular
  uint32_t global_mod_decoded;     // channels fully decoded in the global stream (filled by host plan)
  uint32_t global_nb_transforms;   // frame-level modular transforms (multi-section frames; host-parsed, PlanChannels applied)
  ModTransform global_tr[kMaxTransforms];
  uint32_t global_serial;          // a frame-level palette needs the serial inverse (deltas / predictor): GlobalInverseSerialKernel, not per pixel
  uint32_t num_coded;              // multi-section frames: coded (non-meta) channels once the frame-level palettes are applied
  uint8_t coded_plane[kMaxModPlanes];  // ... and the plane of `mod` each of them is decoded into
  const int32_t* meta;             // palette colours of the frame-level transforms (host-decoded meta channels)
  const int32_t* global_planes;    // multi-section frame no larger than a group: its modular channels [num_coded][height][width], host-decoded
  uint32_t bit_depth;              // bits per sample of the image (implicit palette colours scale with it)
  // ---- planes (device) ----
  int32_t* lf_quant;               // [3][h8][lf_stride]  (Y, X, B as coded)
  uint32_t lf_stride;
  int32_t* xfromy;                 // [h64][w64]
  int32_t* bfromy;
  int32_t* sharpness_i32;          // [h8][lf_stride]
  int32_t* blockinfo;              // per LF group: [2][cells of that LF group]; see blockinfo_off
  const uint32_t* blockinfo_off;   // [num_lf_groups] offset in int32 entries
  uint32_t* nb_blocks;             // [num_lf_groups] decoded block count
  uint32_t* lf_extra_precision;    // [num_lf_groups]
  uint8_t* cell_strategy;          // [h8][w8]: strategy | 0x80 if top-left cell of its block, 0xFF = uncovered
  uint16_t* cell_hfmul;            // [h8][w8]: hf_mul of the covering block
  uint8_t* cell_sharp;             // [h8][w8]
  uint16_t* cell_off;              // [h8][w8]: (dy << 8) | dx offset of the cell from its block's top-left cell
  uint32_t* group_blocks;          // [num_groups][1024]: blocks of each group in raster order of their top-left cell:
                                   //   bx | by << 5 | strategy << 10 | (hf_mul - 1) << 16   (bx, by relative to the group)
  uint32_t* group_nblocks;         // [num_groups]
  uint64_t* group_ac_end_bit;      // [num_groups]: bit position after the AC data (start of the group's modular data)
  int16_t* coef;                   // [3][coef_h][coef_stride] quantised coefficients (X, Y, B), block-rectangle layout
  uint32_t coef_stride, coef_h;
  float* lf;                       // [3][h8][lf_stride] dequantised (X, Y, B)
  uint32_t* large_list;            // [0] = n, [1 .. n] = bx | by << 16: top-left cells of the blocks that no 64x64 region contains
                                   // (reset by LfFinalKernel, filled by ReconRegionTmaKernel, consumed by ReconLargeListKernel)
  float* xyb0;                     // [3][plane_h][plane_stride]
  float* xyb1;
  uint32_t plane_stride, plane_h;
  // frames coded at half resolution (upsampling 2): size of the frame in image pixels and of the upsampled XYB planes
  uint32_t upsampling, up_width, up_height, up_stride, up_h;
  int32_t* frame_bad;              // [1] set by FrameStatusKernel when any entropy-coded section of the frame failed: the
                                   // reconstruction kernels then skip the frame (its metadata planes are garbage)
  int32_t* mod;                    // [num_mod_channels][height][mod_stride]
  // squeezed extra channels (squeeze.h): pyramid channel table, inverse steps, host-decoded global-stream samples, buffer
  const SqChannel* sq_ch;
  const SqStep* sq_steps;
  const int32_t* sq_global_data;
  int32_t* sq_buf;
  uint32_t sq_nch, sq_global, sq_nsteps, sq_pad;
  uint64_t sq_end_bit;             // single-section frames: where the lane resumes after the host-decoded global modular stream
  uint32_t mod_stride;
  int32_t* status;                 // [num_streams] per-stream status (see StreamStatus)
};

}  // namespace jxlb
