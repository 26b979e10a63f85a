#include "plan.h"

#include <cstdlib>

#include <cstring>

#include "vardct_sections.h"

namespace jxlb {

bool UseUnfusedFilters() {
  static const bool v = getenv("JXLB_SIMPLE_FILTERS") != nullptr;
  return v;
}


namespace {
size_t Align(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }
uint32_t RoundUp(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }
}  // namespace

void MakeFramePlan(const ImageMetadata& md, const FrameHeader& fh, const FrameGlobals& g, size_t cs_padded_bytes, FramePlan* plan) {
  FramePlan& p = *plan;
  FrameDev& f = p.proto;
  memset(&f, 0, sizeof f);
  f.width = fh.coded_w;
  f.orientation = md.orientation;
  f.dither_x0 = fh.have_crop ? (uint32_t) (((fh.x0 % 32) + 32) % 32) : 0u;
  f.dither_y0 = fh.have_crop ? (uint32_t) (((fh.y0 % 32) + 32) % 32) : 0u;
  f.height = fh.coded_h;
  f.w8 = (f.width + 7) / 8;
  f.h8 = (f.height + 7) / 8;
  f.w64 = (f.width + 63) / 64;
  f.h64 = (f.height + 63) / 64;
  f.ngx = fh.ngx;
  f.ngy = fh.ngy;
  f.nlfx = fh.nlfx;
  f.nlfy = fh.nlfy;
  f.num_groups = fh.num_groups;
  f.num_lf_groups = fh.num_lf_groups;
  f.num_passes = fh.num_passes;
  f.group_dim = fh.group_dim;
  f.encoding = fh.encoding;
  f.flags = (uint32_t) fh.flags;
  f.single_section = fh.toc_entries == 1;
  f.upsampling = fh.upsampling;
  f.up_width = fh.width;
  f.up_height = fh.height;
  f.x_qm_scale = fh.x_qm_scale;
  f.b_qm_scale = fh.b_qm_scale;
  f.cs_bytes = cs_padded_bytes;
  f.toc_entries = fh.toc_entries;
  for (int c = 0; c < 3; ++c) f.lf_dequant[c] = g.lf_dequant[c];
  f.global_scale = g.global_scale;
  f.quant_lf = g.quant_lf;
  f.bctx = g.bctx;
  f.cfl = g.cfl;
  f.global_tree_nodes = g.tree_nodes;
  f.global_tree_uses_wp = g.tree_uses_wp;
  f.global_tree_max_property = g.tree_max_property;
  f.global_modular_bit = g.global_modular_bit;
  f.num_hf_presets = g.num_hf_presets;
  f.used_orders = g.used_orders;
  f.orders = g.orders;
  f.rf = fh.rf;
  f.num_color_mod_channels = fh.encoding == 1 ? (md.color.color_space == 1 ? 1u : 3u) : 0u;
  f.num_mod_channels = f.num_color_mod_channels + (uint32_t) md.extra.size();
  f.global_mod_decoded = 0;
  f.global_nb_transforms = g.global_mh.nb_transforms;
  for (int i = 0; i < kMaxTransforms; ++i) f.global_tr[i] = g.global_mh.tr[i];
  f.bit_depth = md.bits_per_sample;
  f.num_coded = f.num_mod_channels;
  for (uint32_t i = 0; i < (uint32_t) kMaxModPlanes; ++i) f.coded_plane[i] = (uint8_t) i;
  if (g.global_mh.nb_transforms && !g.squeeze && fh.toc_entries > 1) {
    for (uint32_t t = 0; t < g.global_mh.nb_transforms; ++t)
      if (PaletteNeedsSerialInverse(g.global_mh.tr[t])) f.global_serial = 1;
    f.num_coded = g.chplan.ncoded;
    for (uint32_t i = 0; i < g.chplan.ncoded; ++i) f.coded_plane[i] = g.chplan.coded_plane[i];
  }
  // a channel is decoded in the global stream iff it fits one group in both dimensions (and all earlier ones do)
  if (f.width <= f.group_dim && f.height <= f.group_dim) f.global_mod_decoded = f.num_mod_channels;

  // ---- const region
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o = Align(o + bytes);
    return r;
  };
  p.off_cs = take(cs_padded_bytes);
  p.off_sec_begin = take(fh.toc_entries * sizeof(uint64_t));
  p.off_sec_end = take(fh.toc_entries * sizeof(uint64_t));
  p.off_bctx_map = take(g.bctx_map.size());
  p.off_tree = take(g.tree_blob.size());
  p.off_tree_code = take(g.tree_code.size());
  p.off_ac_code = take(g.ac_code.size());
  p.off_order_pool = take(g.order_pool.size() * sizeof(uint16_t));
  p.off_blockinfo_off = take(fh.num_lf_groups * sizeof(uint32_t));
  if (g.squeeze) {
    f.sq_nch = (uint32_t) g.sq.channels.size();
    f.sq_global = g.sq_global;
    f.sq_nsteps = (uint32_t) g.sq.steps.size();
    f.sq_end_bit = g.sq_end_bit;
    p.off_sq_ch = take(g.sq.channels.size() * sizeof(SqChannel));
    p.off_sq_steps = take(g.sq.steps.size() * sizeof(SqStep));
    p.off_sq_global = take(g.sq_global_data.size() * sizeof(int32_t));
  }
  if (!g.meta_data.empty()) p.off_meta = take(g.meta_data.size() * sizeof(int32_t));
  if (!g.global_planes.empty()) p.off_global_planes = take(g.global_planes.size() * sizeof(int32_t));
  f.pass_shift0 = fh.num_passes > 1 ? fh.pass_shift[0] : 0;
  {
    // which modular channels each pass carries (ISO/IEC 18181-1 passes: the downsampling a pass completes); a single pass
    // carries shifts 0 .. 2, the LF groups everything from 3 up
    int max_shift = 2, min_shift = 3;
    for (uint32_t ps = 0; ps < fh.num_passes && ps < 12; ++ps) {
      for (uint32_t j = 0; j < fh.num_ds && j < 4; ++j)
        if (fh.pass_last[j] == ps) min_shift = fh.pass_downsample[j] == 8 ? 3 : fh.pass_downsample[j] == 4 ? 2 : fh.pass_downsample[j] == 2 ? 1 : 0;
      if (ps + 1 == fh.num_passes) min_shift = 0;
      f.pass_min_shift[ps] = (uint8_t) min_shift;
      f.pass_max_shift[ps] = (uint8_t) (max_shift < 0 ? 0 : max_shift);
      if (max_shift < min_shift) f.pass_min_shift[ps] = 255;  // empty bracket: the pass carries no modular channel
      max_shift = min_shift - 1;
    }
  }
  if (!g.extra_passes.empty()) {
    // [PassDev x (num_passes - 1)] then, per pass, its order pool and its code blob (offsets relative to the table)
    size_t bytes = Align(g.extra_passes.size() * sizeof(PassDev));
    for (const auto& ep : g.extra_passes) bytes += Align(ep.order_pool.size() * 2) + Align(ep.ac_code.size());
    p.off_pass_table = take(bytes);
  }
  p.const_bytes = o;

  // ---- work region
  o = 0;
  const bool vardct = fh.encoding == 0;
  f.lf_stride = RoundUp(f.w8, 32);
  f.coef_stride = RoundUp(f.w8 * 8, 64);
  f.coef_h = f.h8 * 8;
  // XYB planes cover whole 64x64 regions: the inverse-transform kernel stores every region row with one 256-byte bulk copy
  f.plane_stride = RoundUp(f.w8 * 8, 64);
  f.plane_h = RoundUp(f.h8 * 8, 64);
  f.mod_stride = RoundUp(f.width, 32);
  p.num_streams = fh.num_lf_groups + fh.num_groups + 1;
  p.off_status = take(p.num_streams * sizeof(int32_t));
  p.off_frame_bad = take(sizeof(int32_t));
  if (vardct) {
    p.off_lf_quant = take((size_t) 3 * f.h8 * f.lf_stride * 4);
    p.off_xfromy = take((size_t) f.h64 * f.w64 * 4);
    p.off_bfromy = take((size_t) f.h64 * f.w64 * 4);
    p.off_sharp_i32 = take((size_t) f.h8 * f.lf_stride * 4);
    // BlockInfo: per LF group 2 x (cells of the group)
    size_t total_cells = 0;
    for (uint32_t l = 0; l < fh.num_lf_groups; ++l) {
      uint32_t cx0, cy0, w8, h8, w64, h64;
      LfGroupRect(f, l, &cx0, &cy0, &w8, &h8, &w64, &h64);
      total_cells += (size_t) w8 * h8;
    }
    p.off_blockinfo = take(total_cells * 2 * 4);
    p.off_nb_blocks = take(fh.num_lf_groups * 4);
    p.off_extra_prec = take(fh.num_lf_groups * 4);
    p.off_cell_strategy = take((size_t) f.h8 * f.w8);
    p.off_cell_hfmul = take((size_t) f.h8 * f.w8 * 2);
    p.off_cell_sharp = take((size_t) f.h8 * f.w8);
    p.off_cell_off = take((size_t) f.h8 * f.w8 * 2);
    p.off_group_blocks = take((size_t) fh.num_groups * 1024 * 4);
    p.off_group_nblocks = take((size_t) fh.num_groups * 4);
    p.off_group_ac_end = take((size_t) fh.num_groups * 8);
    p.coef_bytes = (size_t) 3 * f.coef_h * f.coef_stride * 2;
    p.off_coef = take(p.coef_bytes);
    p.off_lf = take((size_t) 3 * f.h8 * f.lf_stride * 4);
    p.off_large_list = take(((size_t) f.h8 * f.w8 / 2 + 2) * 4);  // [0] = count, then one entry per block outside the region kernel
    p.xyb_bytes = (size_t) 3 * f.plane_h * f.plane_stride * 4;
    if (fh.upsampling == 2) {
      f.up_stride = RoundUp(2 * f.width, 64);
      f.up_h = 2 * f.height;
      p.up_bytes = (size_t) (3 + (md.extra.empty() ? 0 : 1)) * f.up_h * f.up_stride * 4;  // XYB planes (+ the upsampled alpha plane)
    }
  }
  if (f.num_mod_channels) p.off_mod = take((size_t) f.num_mod_channels * f.height * f.mod_stride * 4);
  if (g.squeeze) p.off_sq_buf = take(g.sq.buffer_ints * sizeof(int32_t));
  p.work_bytes = o;
}

void FillConstRegion(const FramePlan& plan, const uint8_t* cs_padded, const FrameHeader& fh, const FrameGlobals& g, uint8_t* dst) {
  // zero everything but the codestream copy, which is the bulk of the region and overwritten next
  memset(dst, 0, plan.off_cs);
  const size_t cs_end = plan.off_cs + plan.proto.cs_bytes;
  memset(dst + cs_end, 0, plan.const_bytes - cs_end);
  memcpy(dst + plan.off_cs, cs_padded, plan.proto.cs_bytes);
  memcpy(dst + plan.off_sec_begin, fh.sec_bit_begin.data(), fh.sec_bit_begin.size() * 8);
  memcpy(dst + plan.off_sec_end, fh.sec_bit_end.data(), fh.sec_bit_end.size() * 8);
  if (!g.bctx_map.empty()) memcpy(dst + plan.off_bctx_map, g.bctx_map.data(), g.bctx_map.size());
  if (!g.tree_blob.empty()) memcpy(dst + plan.off_tree, g.tree_blob.data(), g.tree_blob.size());
  if (!g.tree_code.empty()) memcpy(dst + plan.off_tree_code, g.tree_code.data(), g.tree_code.size());
  if (!g.ac_code.empty()) memcpy(dst + plan.off_ac_code, g.ac_code.data(), g.ac_code.size());
  if (!g.order_pool.empty()) memcpy(dst + plan.off_order_pool, g.order_pool.data(), g.order_pool.size() * 2);
  if (g.squeeze) {
    memcpy(dst + plan.off_sq_ch, g.sq.channels.data(), g.sq.channels.size() * sizeof(SqChannel));
    memcpy(dst + plan.off_sq_steps, g.sq.steps.data(), g.sq.steps.size() * sizeof(SqStep));
    memcpy(dst + plan.off_sq_global, g.sq_global_data.data(), g.sq_global_data.size() * sizeof(int32_t));
  }
  if (!g.meta_data.empty()) memcpy(dst + plan.off_meta, g.meta_data.data(), g.meta_data.size() * sizeof(int32_t));
  if (!g.global_planes.empty()) memcpy(dst + plan.off_global_planes, g.global_planes.data(), g.global_planes.size() * sizeof(int32_t));
  if (!g.extra_passes.empty()) {
    uint8_t* base = dst + plan.off_pass_table;
    PassDev* pd = reinterpret_cast<PassDev*>(base);
    size_t rel = Align(g.extra_passes.size() * sizeof(PassDev));
    for (size_t i = 0; i < g.extra_passes.size(); ++i) {
      const auto& ep = g.extra_passes[i];
      pd[i].orders = ep.orders;
      pd[i].used_orders = ep.used_orders;
      pd[i].shift = i + 1 < (size_t) fh.num_passes ? fh.pass_shift[i + 1] : 0;
      pd[i].order_pool_rel = rel;
      if (!ep.order_pool.empty()) memcpy(base + rel, ep.order_pool.data(), ep.order_pool.size() * 2);
      rel += Align(ep.order_pool.size() * 2);
      pd[i].ac_code_rel = rel;
      memcpy(base + rel, ep.ac_code.data(), ep.ac_code.size());
      rel += Align(ep.ac_code.size());
    }
  }
  uint32_t* bo = reinterpret_cast<uint32_t*>(dst + plan.off_blockinfo_off);
  size_t acc = 0;
  for (uint32_t l = 0; l < fh.num_lf_groups; ++l) {
    uint32_t cx0, cy0, w8, h8, w64, h64;
    LfGroupRect(plan.proto, l, &cx0, &cy0, &w8, &h8, &w64, &h64);
    bo[l] = (uint32_t) acc;
    acc += (size_t) w8 * h8 * 2;
  }
}

FrameDev BindFrameDev(const FramePlan& p, const uint8_t* cb, uint8_t* wb) {
  FrameDev f = p.proto;
  f.cs = cb + p.off_cs;
  f.sec_bit_begin = reinterpret_cast<const uint64_t*>(cb + p.off_sec_begin);
  f.sec_bit_end = reinterpret_cast<const uint64_t*>(cb + p.off_sec_end);
  f.bctx_map = cb + p.off_bctx_map;
  f.global_tree = f.global_tree_nodes ? reinterpret_cast<const TreeNode*>(cb + p.off_tree) : nullptr;
  f.global_code = f.global_tree_nodes ? cb + p.off_tree_code : nullptr;
  f.ac_code = cb + p.off_ac_code;
  f.order_pool = reinterpret_cast<const uint16_t*>(cb + p.off_order_pool);
  f.blockinfo_off = reinterpret_cast<const uint32_t*>(cb + p.off_blockinfo_off);
  f.status = reinterpret_cast<int32_t*>(wb + p.off_status);
  f.frame_bad = reinterpret_cast<int32_t*>(wb + p.off_frame_bad);
  if (f.encoding == 0) {
    f.lf_quant = reinterpret_cast<int32_t*>(wb + p.off_lf_quant);
    f.xfromy = reinterpret_cast<int32_t*>(wb + p.off_xfromy);
    f.bfromy = reinterpret_cast<int32_t*>(wb + p.off_bfromy);
    f.sharpness_i32 = reinterpret_cast<int32_t*>(wb + p.off_sharp_i32);
    f.blockinfo = reinterpret_cast<int32_t*>(wb + p.off_blockinfo);
    f.large_list = reinterpret_cast<uint32_t*>(wb + p.off_large_list);
    f.nb_blocks = reinterpret_cast<uint32_t*>(wb + p.off_nb_blocks);
    f.lf_extra_precision = reinterpret_cast<uint32_t*>(wb + p.off_extra_prec);
    f.cell_strategy = wb + p.off_cell_strategy;
    f.cell_hfmul = reinterpret_cast<uint16_t*>(wb + p.off_cell_hfmul);
    f.cell_sharp = wb + p.off_cell_sharp;
    f.cell_off = reinterpret_cast<uint16_t*>(wb + p.off_cell_off);
    f.group_blocks = reinterpret_cast<uint32_t*>(wb + p.off_group_blocks);
    f.group_nblocks = reinterpret_cast<uint32_t*>(wb + p.off_group_nblocks);
    f.group_ac_end_bit = reinterpret_cast<uint64_t*>(wb + p.off_group_ac_end);
    f.coef = reinterpret_cast<int16_t*>(wb + p.off_coef);
    f.lf = reinterpret_cast<float*>(wb + p.off_lf);
    f.xyb0 = nullptr;  // bound by the caller (plan.xyb_bytes each)
    f.xyb1 = nullptr;
  }
  if (f.num_mod_channels) f.mod = reinterpret_cast<int32_t*>(wb + p.off_mod);
  if (f.sq_nch) {
    f.sq_ch = reinterpret_cast<const SqChannel*>(cb + p.off_sq_ch);
    f.sq_steps = reinterpret_cast<const SqStep*>(cb + p.off_sq_steps);
    f.sq_global_data = reinterpret_cast<const int32_t*>(cb + p.off_sq_global);
    f.sq_buf = reinterpret_cast<int32_t*>(wb + p.off_sq_buf);
  }
  f.meta = p.off_meta ? reinterpret_cast<const int32_t*>(cb + p.off_meta) : nullptr;
  f.global_planes = p.off_global_planes ? reinterpret_cast<const int32_t*>(cb + p.off_global_planes) : nullptr;
  f.pass_table = p.off_pass_table ? reinterpret_cast<const PassDev*>(cb + p.off_pass_table) : nullptr;
  return f;
}

}  // namespace jxlb
