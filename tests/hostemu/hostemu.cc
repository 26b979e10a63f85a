// TEST TOOL — not part of the product and never loaded by it.
// Runs the *shared host/device* section decoders (jxl_coder_b200/csrc/{entropy,modular,vardct_sections}.h) serially on
// the CPU over host buffers laid out by the same FramePlan the GPU decoder uses.  tests/test_sections_host.py compares
// the resulting planes with the Python oracle, so the bit-level logic of the CUDA kernels is verified on a box without
// a GPU; the GPU tests then only need to show that the kernels produce the same planes as this run.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#if defined(__x86_64__) || defined(__i386__)
#include <xmmintrin.h>
#endif
#include "../../jxl_coder_b200/csrc/resize.h"
#include "../../jxl_coder_b200/csrc/color_matrix.h"

#include "../../jxl_coder_b200/csrc/frame_parser.h"
#include "../../jxl_coder_b200/csrc/plan.h"
#include "../../jxl_coder_b200/csrc/test_block.h"
#include "../../jxl_coder_b200/csrc/vardct_sections.h"
#include "../../jxl_coder_b200/csrc/color_params.h"
#include "../../jxl_coder_b200/csrc/numeric_tables.h"
#include "../../jxl_coder_b200/csrc/pixel_stages.h"
#include "../../jxl_coder_b200/csrc/recon.h"

using namespace jxlb;

struct Emu {
  ImageMetadata md;
  FrameHeader fh;
  FrameGlobals g;
  FramePlan plan;
  ByteVec cs;
  std::vector<uint8_t> cregion, wregion;
  std::vector<float> xyb;
  FrameDev f;
  int status = 0;
  int failed_stream = -1;
};

static void SetErr(char* err, size_t n, const std::string& s) {
  if (err && n) snprintf(err, n, "%s", s.c_str());
}

extern "C" {

// frame_index: which frame of the codestream to decode (0 = first).
void* emu_decode(const uint8_t* data, size_t len, int frame_index, char* err, size_t errlen) {
  Emu* e = new Emu();
  size_t cs_len = 0;
  std::string msg;
  int st = ExtractCodestream(data, len, &e->cs, &cs_len);
  if (st) {
    SetErr(err, errlen, "extract codestream: " + std::to_string(st));
    delete e;
    return nullptr;
  }
  uint64_t frame_bit = 0;
  st = ParseImageHeader(e->cs.data(), e->cs.size(), cs_len, &e->md, &frame_bit, &msg);
  if (st) {
    SetErr(err, errlen, "image header: " + msg);
    delete e;
    return nullptr;
  }
  for (int i = 0;; ++i) {
    e->fh = FrameHeader();
    st = ParseFrameHeader(e->cs.data(), e->cs.size(), cs_len, e->md, frame_bit, &e->fh, &msg);
    if (st) {
      SetErr(err, errlen, "frame header: " + msg);
      delete e;
      return nullptr;
    }
    if (i == frame_index) break;
    if (e->fh.is_last) {
      SetErr(err, errlen, "no such frame");
      delete e;
      return nullptr;
    }
    frame_bit = e->fh.end_byte * 8;
  }
  st = ParseFrameGlobals(e->cs.data(), e->cs.size(), e->md, e->fh, &e->g, &msg);
  if (st) {
    SetErr(err, errlen, "frame globals (" + std::to_string(st) + "): " + msg);
    delete e;
    return nullptr;
  }
  MakeFramePlan(e->md, e->fh, e->g, e->cs.size(), &e->plan);
  e->cregion.resize(e->plan.const_bytes);
  e->wregion.assign(e->plan.work_bytes, 0);
  FillConstRegion(e->plan, e->cs.data(), e->fh, e->g, e->cregion.data());
  e->f = BindFrameDev(e->plan, e->cregion.data(), e->wregion.data());
  e->xyb.assign(e->plan.xyb_bytes / sizeof(float), 0.0f);
  e->f.xyb0 = e->xyb.data();
  const FrameDev& f = e->f;
  // scratch
  std::vector<uint8_t> arena_mem(8u << 20), hf_mem(8u << 20);
  std::vector<int32_t> wp(ModFastScratch::Ints(65536 + 64));
  std::vector<uint32_t> lz(1u << 20), perm(2 * 65536);
  std::vector<uint8_t> nz(3 * 1024);
  StreamScratch s;
  s.arena.Init(arena_mem.data(), (uint32_t) arena_mem.size());
  s.wp = wp.data();
  s.wp_ints = (uint32_t) wp.size();
  s.lz77 = lz.data();
  s.lz77_mask = (1u << 20) - 1;
  if (getenv("EMU_NO_LZ77")) {  // what the device lanes have: no LZ77 window
    s.lz77 = nullptr;
    s.lz77_mask = 0;
  }
  s.nzmap = nz.data();
  const NaturalOrders& nat = NaturalOrderPoolHost();
  const uint32_t kMaxNodes = 1u << 16;
  if (f.single_section) {
    Arena hf;
    hf.Init(hf_mem.data(), (uint32_t) hf_mem.size());
    e->status = DecodeSingleSectionFrame(f, nat, s, hf, perm.data(), kMaxNodes);
    if (e->status) e->failed_stream = 0;
    if (!e->status && f.sq_nch) UnsqueezeAllSerial(f);
  } else {
    for (uint32_t l = 0; l < f.num_lf_groups && !e->status && f.encoding == 0; ++l) {
      BitReader br;
      br.Init(f.cs, f.cs_bytes, f.sec_bit_begin[1 + l], f.sec_bit_end[1 + l]);
      s.arena.used = 0;
      e->status = DecodeLfGroupSection(br, f, l, s, kMaxNodes);
      if (e->status) e->failed_stream = (int) l;
    }
    for (uint32_t gi = 0; gi < f.num_groups && !e->status; ++gi) {
      for (uint32_t pass = 0; pass < f.num_passes && !e->status; ++pass) {
        uint32_t sec = 1 + f.num_lf_groups + 1 + pass * f.num_groups + gi;
        BitReader br;
        br.Init(f.cs, f.cs_bytes, f.sec_bit_begin[sec], f.sec_bit_end[sec]);
        s.arena.used = 0;
        if (f.encoding == 0) e->status = DecodeAcGroup(br, f, gi, nat, s, pass);
        if (!e->status) e->status = DecodeModularGroup(br, f, gi, s, kMaxNodes, pass);
      }
      if (e->status) e->failed_stream = (int) (f.num_lf_groups + gi);
    }
    if (!e->status && f.sq_nch) UnsqueezeAllSerial(f);
    if (!e->status && f.global_planes)  // ScatterGlobalPlanesKernel
      for (uint32_t c = 0; c < f.num_coded; ++c)
        for (uint32_t y = 0; y < f.height; ++y)
          for (uint32_t x = 0; x < f.width; ++x)
            f.mod[(size_t) f.coded_plane[c] * f.height * f.mod_stride + (size_t) y * f.mod_stride + x] = f.global_planes[((size_t) c * f.height + y) * f.width + x];
    // frame-level transforms on the extra channels of a VarDCT frame (the device runs ModularGlobalInverseKernel)
    if (!e->status && !f.sq_nch && f.num_mod_channels && f.global_serial) {  // delta palette: the serial inverse (GlobalInverseSerialKernel)
      ModChannel planes[kMaxModPlanes];
      for (uint32_t c = 0; c < f.num_mod_channels; ++c) planes[c] = ModChannel{f.mod + (size_t) c * f.height * f.mod_stride, f.width, f.height, f.mod_stride};
      e->status = ApplyInverseTransforms(f.global_tr, f.global_nb_transforms, planes, f.meta, f.bit_depth);
    } else if (!e->status && f.encoding == 0 && !f.sq_nch && f.num_mod_channels && f.global_nb_transforms)
      for (uint32_t y = 0; y < f.height; ++y)
        for (uint32_t x = 0; x < f.width; ++x) StageGlobalInverse(f, (int) x, (int) y);
  }
  return e;
}

void emu_free(void* h) { delete static_cast<Emu*>(h); }
int emu_status(void* h) { return static_cast<Emu*>(h)->status; }
int emu_failed_stream(void* h) { return static_cast<Emu*>(h)->failed_stream; }
// Status a pixel stage raised while rendering (a palette pixel it cannot reconstruct: pixel_stages.h StageGlobalInverse).
int emu_late_status(void* h) {
  Emu* e = static_cast<Emu*>(h);
  return e->f.single_section ? 0 : e->f.status[e->f.num_lf_groups];
}

// out[]: width, height, w8, h8, lf_stride, coef_stride, coef_h, num_groups, num_lf_groups, encoding, num_mod_channels,
//        mod_stride, w64, h64, single_section, global_nb_transforms, xsize, ysize, bits, num_extra, is_last, orientation,
//        upsampling, up_width, up_height
void emu_info(void* h, uint32_t* out) {
  Emu* e = static_cast<Emu*>(h);
  const FrameDev& f = e->f;
  uint32_t v[] = {f.width, f.height, f.w8, f.h8, f.lf_stride, f.coef_stride, f.coef_h, f.num_groups, f.num_lf_groups, f.encoding,
                  f.num_mod_channels, f.mod_stride, f.w64, f.h64, f.single_section, f.global_nb_transforms, e->md.xsize, e->md.ysize,
                  e->md.bits_per_sample, (uint32_t) e->md.extra.size(), (uint32_t) e->fh.is_last, e->md.orientation,
                  f.upsampling, f.up_width, f.up_height};
  memcpy(out, v, sizeof v);
}
const int32_t* emu_lf_quant(void* h) { return static_cast<Emu*>(h)->f.lf_quant; }
const int32_t* emu_xfromy(void* h) { return static_cast<Emu*>(h)->f.xfromy; }
const int32_t* emu_bfromy(void* h) { return static_cast<Emu*>(h)->f.bfromy; }
const uint8_t* emu_cell_strategy(void* h) { return static_cast<Emu*>(h)->f.cell_strategy; }
const uint16_t* emu_cell_hfmul(void* h) { return static_cast<Emu*>(h)->f.cell_hfmul; }
const uint8_t* emu_cell_sharp(void* h) { return static_cast<Emu*>(h)->f.cell_sharp; }
const int16_t* emu_coef(void* h) { return static_cast<Emu*>(h)->f.coef; }
const int32_t* emu_mod(void* h) { return static_cast<Emu*>(h)->f.mod; }

// Runs the numeric stages (the same host/device functions the kernels call) and writes RGBA8 / RGBA16.
// also exposes the XYB planes after the inverse transforms (stage = 0) and after the filters (stage = 1).
int emu_render(void* h, uint8_t* out, uint32_t stride_bytes, int bits16, float* xyb_idct, float* xyb_final) {
  Emu* e = static_cast<Emu*>(h);
  FrameDev& f = e->f;
  const NumericTables& nt = GetHostNumericTables().tables;
  OutputDesc od;
  od.data = out;
  od.stride_bytes = stride_bytes;
  od.bits16 = (uint32_t) bits16;
  od.alpha_channel = -1;
  od.alpha_bits = 8;
  od.color_bits = e->md.bits_per_sample;
  int ai = e->md.alpha_channel();
  if (ai >= 0) {
    od.alpha_channel = (int32_t) (f.num_color_mod_channels + (uint32_t) ai);
    od.alpha_bits = e->md.extra[ai].bits;
  }
  auto nosync = [] {};
  if (f.encoding == 0) {
    ColorParams cp;
    std::string err;
    if (MakeColorParams(e->md, &cp, &err)) return 3;
    const LfMul m = MakeLfMul(f);
    const size_t lfplane = (size_t) f.h8 * f.lf_stride;
    for (uint32_t cy = 0; cy < f.h8; ++cy)
      for (uint32_t cx = 0; cx < f.w8; ++cx) {
        float v[3];
        LfFinalCell(f, m, cx, cy, v);
        for (int c = 0; c < 3; ++c) f.lf[c * lfplane + (size_t) cy * f.lf_stride + cx] = v[c];
      }
    static RegionShared sh;
    for (uint32_t ry = 0; ry < (f.h8 + 7) / 8; ++ry)
      for (uint32_t rx = 0; rx < (f.w8 + 7) / 8; ++rx) ReconRegion(f, nt, rx, ry, sh, 0, 1, nosync);
    for (uint32_t cy = 0; cy < f.h8; ++cy)
      for (uint32_t cx = 0; cx < f.w8; ++cx) {
        uint8_t s = f.cell_strategy[(size_t) cy * f.w8 + cx];
        if ((s & 0x80) && s != 0xFF && BlockNeedsLargePath(s & 0x7F, cx, cy)) ReconLargeBlock(f, nt, cx, cy, 0, 1, nosync);
      }
    const size_t plane = (size_t) f.plane_h * f.plane_stride;
    if (xyb_idct) memcpy(xyb_idct, f.xyb0, 3 * plane * sizeof(float));
    float* src = f.xyb0;
    std::vector<float> second(3 * plane);  // the plan only carries a second XYB buffer for the unfused debug kernels
    float* dst = second.data();
    auto run = [&](auto fn) {
      for (uint32_t y = 0; y < f.height; ++y)
        for (uint32_t x = 0; x < f.width; ++x) fn((int) x, (int) y);
      float* t = src;
      src = dst;
      dst = t;
    };
    if (f.rf.gab) run([&](int x, int y) { StageGaborish(f, src, dst, x, y); });
    if (f.rf.epf_iters == 3) run([&](int x, int y) { StageEpf(f, nt, 0, src, dst, x, y); });
    if (f.rf.epf_iters >= 1) run([&](int x, int y) { StageEpf(f, nt, 1, src, dst, x, y); });
    if (f.rf.epf_iters >= 2) run([&](int x, int y) { StageEpf(f, nt, 2, src, dst, x, y); });
    if (xyb_final) memcpy(xyb_final, src, 3 * plane * sizeof(float));
    if (f.upsampling == 2) {
      std::vector<float> up((size_t) 3 * f.up_h * f.up_stride);
      for (uint32_t y = 0; y < f.height; ++y)
        for (uint32_t x = 0; x < f.width; ++x) StageUpsample2(f, src, up.data(), f.up_stride, f.up_h, (int) x, (int) y);
      FrameDev fu = f;
      fu.width = f.up_width;
      fu.height = f.up_height;
      fu.plane_stride = f.up_stride;
      fu.plane_h = f.up_h;
      OutputDesc odu = od;
      std::vector<int32_t> up_alpha;
      if (od.alpha_channel >= 0) {
        up_alpha.resize((size_t) f.up_h * f.up_stride);
        const int32_t* a = f.mod + (size_t) od.alpha_channel * f.height * f.mod_stride;
        for (uint32_t y = 0; y < f.height; ++y)
          for (uint32_t x = 0; x < f.width; ++x)
            StageUpsampleAlpha2(f, a, od.alpha_bits, up_alpha.data(), f.up_stride, (int) x, (int) y);
        fu.mod = up_alpha.data();
        fu.mod_stride = f.up_stride;
        odu.alpha_channel = 0;
        odu.alpha_float = 1;
      }
      for (uint32_t y = 0; y < fu.height; ++y)
        for (uint32_t x = 0; x < fu.width; ++x) StageColorToRgba(fu, cp, nt, up.data(), odu, (int) x, (int) y);
      return 0;
    }
    for (uint32_t y = 0; y < f.height; ++y)
      for (uint32_t x = 0; x < f.width; ++x) StageColorToRgba(f, cp, nt, src, od, (int) x, (int) y);
  } else {
    if (e->md.xyb_encoded) return 3;
    for (uint32_t y = 0; y < f.height; ++y)
      for (uint32_t x = 0; x < f.width; ++x) {
        if (!f.single_section && !f.global_serial) StageGlobalInverse(f, (int) x, (int) y);
        StageModularToRgba(f, od, (int) x, (int) y);
      }
  }
  return 0;
}
uint32_t emu_plane_stride(void* h) { return static_cast<Emu*>(h)->f.plane_stride; }
uint32_t emu_plane_h(void* h) { return static_cast<Emu*>(h)->f.plane_h; }

// Unit hooks for table checks.
uint32_t emu_freq_ctx(uint32_t k) { return ZeroDensityFreqCtx(k); }
uint32_t emu_nnz_ctx(uint32_t k) { return ZeroDensityNnzCtx(k); }
// One synthetic block through the reconstruction code (test_block.h): coeffs_out = the dequantised + LLF coefficient arrays
// F[c][vertical][horizontal] the inverse transform consumes, pixels_out = its result; both [3][8 cy][8 cx].
int emu_recon_block(uint32_t strategy, const int16_t* q, const float* lf, uint32_t hf_mul, uint32_t global_scale, float* coeffs_out,
                    float* pixels_out) {
  // plain DCT strategies only: the special 8x8 transforms (IDENTITY, DCT2X2, DCT4X4, DCT4X8, AFV) keep their own coefficient
  // layouts and are reached by encoder-made files
  if (strategy >= (uint32_t) kNumStrategies || (strategy >= 1 && strategy <= 3) || (strategy >= 12 && strategy <= 17)) return 1;
  ImageMetadata md;
  FrameHeader fh;
  FrameGlobals g;
  FramePlan plan;
  MakeSingleBlockPlan(strategy, global_scale, &md, &fh, &g, &plan);
  std::vector<uint8_t> cb(plan.const_bytes), wb(plan.work_bytes), cs(16);
  FillConstRegion(plan, cs.data(), fh, g, cb.data());
  FrameDev f = BindFrameDev(plan, cb.data(), wb.data());
  std::vector<float> xyb(plan.xyb_bytes / sizeof(float));
  f.xyb0 = f.xyb1 = xyb.data();
  FillSingleBlock(f, strategy, q, lf, hf_mul);
  const NumericTables& nt = GetHostNumericTables().tables;
  auto nosync = [] {};
  const uint32_t R = 8 * StrategyCellsY(strategy), C = 8 * StrategyCellsX(strategy);
  const size_t pplane = (size_t) f.plane_h * f.plane_stride;
  auto grab = [&](float* dst) {
    for (uint32_t c = 0; c < 3; ++c)
      for (uint32_t r = 0; r < R; ++r) memcpy(dst + ((size_t) c * R + r) * C, f.xyb0 + c * pplane + (size_t) r * f.plane_stride, C * sizeof(float));
  };
  ReconLargeCoefficients(f, nt, 0, 0, 0, 1, nosync);
  grab(coeffs_out);
  ReconLargeInverse(f, 0, 0, 0, 1, nosync);
  grab(pixels_out);
  return 0;
}
// PlanChannels (modular.h) on a hand-made header.  tr: per transform (id, begin_c, num_c, nb_colours).
// out: [0] nb_meta, [1] ncoded, [2..9] coded_plane, [10..17] meta_tr, [18] meta_ints
int emu_plan_channels(uint32_t nfinal, uint32_t ntr, const uint32_t* tr, uint32_t* out) {
  ModularHeader mh{};
  mh.nb_transforms = (uint8_t) ntr;
  for (uint32_t t = 0; t < ntr; ++t) {
    mh.tr[t].id = (uint8_t) tr[4 * t];
    mh.tr[t].begin_c = tr[4 * t + 1];
    mh.tr[t].num_c = tr[4 * t + 2];
    mh.tr[t].nb_colours = tr[4 * t + 3];
  }
  ChannelPlan cp{};
  const int st = PlanChannels(&mh, nfinal, &cp);
  if (st != kOk) return st;
  out[0] = cp.nb_meta;
  out[1] = cp.ncoded;
  for (int i = 0; i < 8; ++i) out[2 + i] = cp.coded_plane[i];
  for (int i = 0; i < 8; ++i) out[10 + i] = cp.meta_tr[i];
  out[18] = cp.meta_ints;
  return 0;
}
// Lehmer digits -> permutation (entropy.h ExpandLehmer); digits is overwritten.
int emu_expand_lehmer(uint32_t size, uint32_t end, uint32_t* perm, uint32_t* digits) { return ExpandLehmer(size, end, perm, digits); }
void emu_logcount(uint32_t idx7, uint32_t* nb, uint32_t* sym) { LogCountLookup(idx7, nb, sym); }
uint32_t emu_natural_order(uint32_t order_id, uint16_t* out) {
  const NaturalOrders& nat = NaturalOrderPoolHost();
  memcpy(out, nat.pool + nat.offset[order_id], nat.size[order_id] * 2);
  return nat.size[order_id];
}


// Rescale (resize.h) of an RGBA8 image on the CPU: returns 0 and fills out (caller provides >= req capacity) or a
// kResize* status.  dims[0..1] = output width / height.
int emu_resize(const uint8_t* src, uint32_t w, uint32_t h, int32_t req_w, int32_t req_h, int32_t scale_mode, int32_t filter, int32_t has_alpha,
               uint8_t* out, size_t out_capacity, uint32_t* dims) {
  ResizePlan plan;
  int st = MakeResizePlan(w, h, req_w, req_h, scale_mode, filter, has_alpha != 0, &plan);
  if (st) return st;
  std::vector<uint8_t> res;
  ResizeRgba8Host(plan, src, w * 4, &res);
  dims[0] = plan.out_w;
  dims[1] = plan.out_h;
  if (res.size() > out_capacity) return -1;
  memcpy(out, res.data(), res.size());
  return 0;
}


// api_level < 34 colour pass (color_matrix.h) of the image whose file is `jxl`, in place on an RGBA8 array.
// 0 applied, 1 not needed for this colour encoding, 2 needed but not covered, -1 parse error.
int emu_color_matrix(const uint8_t* jxl, size_t len, uint8_t* rgba, uint32_t w, uint32_t h) {
  ByteVec cs;
  size_t cs_len = 0;
  if (ExtractCodestream(jxl, len, &cs, &cs_len)) return -1;
  ImageMetadata md;
  uint64_t fb = 0;
  std::string err;
  if (ParseImageHeader(cs.data(), cs.size(), cs_len, &md, &fb, &err)) return -1;
  bool needed = false;
  static ColorMatrixPlan plan;
  if (!MakeColorMatrixPlan(md, &needed, &plan)) return 2;
  if (!needed) return 1;
  ApplyColorMatrixHost(plan, rgba, w * 4, w, h);
  return 0;
}
// The 16-bit variant (applyColorMatrix16Bit) on RGBA16 rows.
int emu_color_matrix16(const uint8_t* jxl, size_t len, uint16_t* rgba, uint32_t w, uint32_t h) {
  ByteVec cs;
  size_t cs_len = 0;
  if (ExtractCodestream(jxl, len, &cs, &cs_len)) return -1;
  ImageMetadata md;
  uint64_t fb = 0;
  std::string err;
  if (ParseImageHeader(cs.data(), cs.size(), cs_len, &md, &fb, &err)) return -1;
  bool needed = false;
  static ColorMatrixPlan plan;
  static ColorMatrixTables16 t16;
  if (!MakeColorMatrixPlan(md, &needed, &plan, &t16)) return 2;
  if (!needed) return 1;
  ApplyColorMatrixHost16(plan, t16, rgba, w * 8, w, h);
  return 0;
}


// ApproxRcp (numeric.h) against the host's RCPPS on n pseudo-random positive normal floats in [2^-20, 2^20): returns the
// number of inputs whose results differ.
long emu_rcp_check(long n) {
#if defined(__x86_64__) || defined(__i386__)
  const NumericTables& nt = GetHostNumericTables().tables;
  long bad = 0;
  uint32_t s = 12345u;
  for (long i = 0; i < n; ++i) {
    s = s * 1664525u + 1013904223u;
    union { float f; uint32_t u; } v;
    v.u = ((107u + (s >> 27) + ((s >> 7) & 7u)) << 23) | (s & 0x7FFFFFu);
    alignas(16) float in4[4] = {v.f, v.f, v.f, v.f}, out4[4];
    _mm_store_ps(out4, _mm_rcp_ps(_mm_load_ps(in4)));
    if (out4[0] != ApproxRcp(nt.rcp11, v.f)) ++bad;
  }
  return bad;
#else
  (void) n;
  return 0;
#endif
}

}  // extern "C"
