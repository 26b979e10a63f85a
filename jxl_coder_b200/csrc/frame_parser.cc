// Host-side header parsing; see frame_parser.h.
#include <algorithm>
#include "frame_parser.h"
#include "modular_fast.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>

#include "vardct_sections.h"

namespace jxlb {

namespace {

const uint8_t kContainerSig[12] = {0, 0, 0, 0xC, 'J', 'X', 'L', ' ', 0xD, 0xA, 0x87, 0xA};

uint64_t BE(const uint8_t* p, int n) {
  uint64_t v = 0;
  for (int i = 0; i < n; ++i) v = (v << 8) | p[i];
  return v;
}

#define JXLB_FAIL(code, msg) \
  do {                       \
    if (err) *err = (msg);   \
    return (code);           \
  } while (0)

// Parse scratch (arenas, permutation buffers, LZ77 windows) comes from a process-wide pool of grow-only blocks whose
// contents are unspecified: a fresh zero-filled 8 MB vector per image costs more than the parse itself (page faults and
// the kernel's per-process mapping lock, which made 8 parse threads 3.5x slower per image than one).
class ScratchLease {
 public:
  explicit ScratchLease(size_t bytes) {
    {
      std::lock_guard<std::mutex> l(Mu());
      auto& fl = Free();
      for (size_t k = 0; k < fl.size(); ++k)
        if (fl[k].cap >= bytes) {
          b_ = std::move(fl[k]);
          fl[k] = std::move(fl.back());
          fl.pop_back();
          Poison();
          return;
        }
    }
    b_.mem.reset(new uint8_t[bytes]);
    b_.cap = bytes;
    Poison();
  }
  ~ScratchLease() {
    std::lock_guard<std::mutex> l(Mu());
    if (Free().size() < 256) Free().push_back(std::move(b_));
  }
  ScratchLease(const ScratchLease&) = delete;
  ScratchLease& operator=(const ScratchLease&) = delete;
  // JXLB_POISON_SCRATCH=1 (debugging aid): every lease starts as 0xCD bytes, so that a read of scratch that was never written
  // shows up as a reproducible difference instead of depending on what the previous image left behind
  void Poison() {
    static const bool on = getenv("JXLB_POISON_SCRATCH") != nullptr;
    if (on) memset(b_.mem.get(), 0xCD, b_.cap);
  }
  uint8_t* data() { return b_.mem.get(); }
  template <class T>
  T* as() { return reinterpret_cast<T*>(b_.mem.get()); }

 private:
  struct Block {
    std::unique_ptr<uint8_t[]> mem;
    size_t cap = 0;
  };
  static std::mutex& Mu() {
    static std::mutex m;
    return m;
  }
  static std::vector<Block>& Free() {
    static std::vector<Block> v;
    return v;
  }
  Block b_;
};

void ReadSizeHeader(BitReader& br, uint32_t* xs, uint32_t* ys) {
  uint32_t small = br.Read(1);
  auto dim = [&]() -> uint32_t { return small ? (br.Read(5) + 1) * 8 : br.U32(1, 9, 1, 13, 1, 18, 1, 30); };
  uint32_t h = dim();
  uint32_t ratio = br.Read(3);
  uint32_t w;
  switch (ratio) {
    case 0: w = dim(); break;
    case 1: w = h; break;
    case 2: w = (uint32_t) ((uint64_t) h * 12 / 10); break;
    case 3: w = (uint32_t) ((uint64_t) h * 4 / 3); break;
    case 4: w = (uint32_t) ((uint64_t) h * 3 / 2); break;
    case 5: w = (uint32_t) ((uint64_t) h * 16 / 9); break;
    case 6: w = (uint32_t) ((uint64_t) h * 5 / 4); break;
    default: w = h * 2; break;
  }
  *xs = w;
  *ys = h;
}

void ReadBitDepth(BitReader& br, bool* is_float, uint32_t* bits, uint32_t* exp_bits) {
  *is_float = br.Read(1);
  if (!*is_float) {
    *bits = br.U32(8, 0, 10, 0, 12, 0, 1, 6);
    *exp_bits = 0;
  } else {
    *bits = br.U32(32, 0, 16, 0, 24, 0, 1, 6);
    *exp_bits = 1 + br.Read(4);
  }
}

int32_t ReadCustomXY(BitReader& br) { return UnpackSigned(br.U32(0, 19, 524288, 19, 1048576, 20, 2097152, 21)); }

}  // namespace

int ExtractCodestream(const uint8_t* data, size_t len, ByteVec* out, size_t* cs_len) {
  out->clear();
  out->reserve(len + 20);  // one allocation: the padding below must not reallocate
  if (len >= 2 && data[0] == 0xFF && data[1] == 0x0A) {
    out->assign(data, data + len);
  } else if (len >= 12 && memcmp(data, kContainerSig, 12) == 0) {
    size_t pos = 0;
    bool any = false;
    while (pos + 8 <= len) {
      uint64_t size = BE(data + pos, 4);
      const uint8_t* typ = data + pos + 4;
      size_t hdr = 8;
      if (size == 1) {
        if (pos + 16 > len) return kParseInvalid;
        size = BE(data + pos + 8, 8);
        hdr = 16;
      } else if (size == 0) {
        size = len - pos;
      }
      if (size < hdr || size > len - pos) {
        // truncated last box: take what is there (the reference then fails with NEED_MORE_INPUT)
        size = len - pos;
        if (size < hdr) break;
      }
      const uint8_t* body = data + pos + hdr;
      size_t blen = (size_t) size - hdr;
      if (memcmp(typ, "jxlc", 4) == 0) {
        out->append(body, body + blen);
        any = true;
      } else if (memcmp(typ, "jxlp", 4) == 0) {
        if (blen < 4) return kParseInvalid;
        out->append(body + 4, body + blen);
        any = true;
      }
      pos += (size_t) size;
    }
    if (!any) return kParseInvalid;
  } else {
    return kParseNotJxl;
  }
  *cs_len = out->size();
  size_t padded = ((out->size() + 3) & ~(size_t) 3) + 16;
  out->resize(padded, (uint8_t) 0);
  return kParseOk;
}

int ParseImageHeader(const uint8_t* cs, size_t cs_padded, size_t cs_len, ImageMetadata* md, uint64_t* frame_bit, std::string* err) {
  if (cs_len < 2 || cs[0] != 0xFF || cs[1] != 0x0A) JXLB_FAIL(kParseNotJxl, "bad codestream signature");
  BitReader br;
  br.Init(cs, cs_padded, 16, (uint64_t) cs_len * 8);
  ReadSizeHeader(br, &md->xsize, &md->ysize);
  bool all_default = br.Read(1);
  bool extra_fields = false;
  if (!all_default) {
    extra_fields = br.Read(1);
    if (extra_fields) {
      md->orientation = 1 + br.Read(3);
      md->have_intrinsic_size = br.Read(1);
      if (md->have_intrinsic_size) {
        uint32_t a, b;
        ReadSizeHeader(br, &a, &b);
      }
      md->have_preview = br.Read(1);
      if (md->have_preview) {
        bool div8 = br.Read(1);
        if (div8) br.U32(16, 0, 32, 0, 1, 5, 33, 9);
        else br.U32(1, 6, 65, 8, 321, 10, 1345, 12);
        uint32_t ratio = br.Read(3);
        if (ratio == 0) {
          if (div8) br.U32(16, 0, 32, 0, 1, 5, 33, 9);
          else br.U32(1, 6, 65, 8, 321, 10, 1345, 12);
        }
      }
      md->have_animation = br.Read(1);
      if (md->have_animation) {
        md->tps_num = br.U32(100, 0, 1000, 0, 1, 10, 1, 30);
        md->tps_den = br.U32(1, 0, 1001, 0, 1, 8, 1, 10);
        md->num_loops = br.U32(0, 0, 0, 3, 0, 16, 0, 32);
        md->have_timecodes = br.Read(1);
      }
    }
    ReadBitDepth(br, &md->float_samples, &md->bits_per_sample, &md->exp_bits);
    md->modular_16bit = br.Read(1);
    uint32_t nextra = br.U32(0, 0, 1, 0, 2, 4, 1, 12);
    md->extra.resize(nextra);
    for (uint32_t i = 0; i < nextra; ++i) {
      ExtraChannelInfo& ec = md->extra[i];
      bool d_alpha = br.Read(1);
      if (!d_alpha) {
        ec.type = br.Enum();
        ReadBitDepth(br, &ec.is_float, &ec.bits, &ec.exp_bits);
        ec.dim_shift = br.U32(0, 0, 3, 0, 4, 0, 1, 3);
        uint32_t name_len = br.U32(0, 0, 0, 4, 16, 5, 48, 10);
        for (uint32_t k = 0; k < name_len; ++k) br.Read(8);
        if (ec.type == 0) {
          ec.alpha_premultiplied = br.Read(1);
        } else if (ec.type == 2) {  // spot colour
          for (int k = 0; k < 4; ++k) br.F16();
        } else if (ec.type == 5) {  // CFA
          br.U32(1, 0, 0, 2, 3, 4, 19, 8);
        }
      }
    }
    md->xyb_encoded = br.Read(1);
    ColorEncoding& ce = md->color;
    ce.all_default = br.Read(1);
    if (!ce.all_default) {
      ce.want_icc = br.Read(1);
      ce.color_space = br.Enum();
      if (!ce.want_icc) {
        if (ce.color_space != 2) {
          ce.white_point = br.Enum();
          if (ce.white_point == 2) {
            ce.white_xy[0] = ReadCustomXY(br);
            ce.white_xy[1] = ReadCustomXY(br);
          }
        }
        if (ce.color_space != 1 && ce.color_space != 2) {
          ce.primaries = br.Enum();
          if (ce.primaries == 2)
            for (int k = 0; k < 3; ++k) {
              ce.prim_xy[k][0] = ReadCustomXY(br);
              ce.prim_xy[k][1] = ReadCustomXY(br);
            }
        }
        ce.have_gamma = br.Read(1);
        if (ce.have_gamma) {
          ce.gamma_u24 = br.Read(24);
          ce.transfer = 0xFFFF;
        } else {
          ce.transfer = br.Enum();
        }
        ce.rendering_intent = br.Enum();
      }
    }
    if (extra_fields) {
      bool tm_default = br.Read(1);
      if (!tm_default) {
        md->intensity_target = br.F16();
        md->min_nits = br.F16();
        md->relative_to_max_display = br.Read(1);
        md->linear_below = br.F16();
      }
    }
    uint64_t ext = br.U64();
    if (ext != 0) JXLB_FAIL(kParseUnsupported, "image metadata extensions");
  }
  md->default_transform = br.Read(1);
  if (!md->default_transform) {
    if (md->xyb_encoded) {
      bool opsin_default = br.Read(1);
      if (!opsin_default) {
        for (int i = 0; i < 9; ++i) md->opsin_inverse[i] = br.F16();
        for (int i = 0; i < 3; ++i) md->opsin_bias[i] = br.F16();
        for (int i = 0; i < 3; ++i) md->quant_bias[i] = br.F16();
        md->quant_bias_numerator = br.F16();
        JXLB_FAIL(kParseUnsupported, "custom opsin inverse matrix");
      }
    }
    uint32_t cw_mask = br.Read(3);
    if (cw_mask) {
      md->custom_upsampling = true;
      if (cw_mask & 1) for (int i = 0; i < 15; ++i) br.F16();
      if (cw_mask & 2) for (int i = 0; i < 55; ++i) br.F16();
      if (cw_mask & 4) for (int i = 0; i < 210; ++i) br.F16();
    }
  }
  if (md->color.want_icc) JXLB_FAIL(kParseUnsupported, "embedded ICC profile in codestream");
  // a preview frame (own size, own frame-header layout) would precede the first frame: not handled, and parsing it as
  // the main frame would give a corrupt-section error or a wrong picture
  if (md->have_preview) JXLB_FAIL(kParseUnsupported, "preview frame");
  if (br.Overrun()) JXLB_FAIL(kParseInvalid, "truncated image header");
  br.AlignToByte();
  *frame_bit = br.Position();
  if (md->xsize == 0 || md->ysize == 0) JXLB_FAIL(kParseInvalid, "empty image");
  return kParseOk;
}

int ParseFrameHeader(const uint8_t* cs, size_t cs_padded, size_t cs_len, const ImageMetadata& md, uint64_t frame_bit,
                     FrameHeader* fh, std::string* err) {
  BitReader br;
  br.Init(cs, cs_padded, frame_bit, (uint64_t) cs_len * 8);
  const uint32_t nextra = (uint32_t) md.extra.size();
  fh->width = md.xsize;
  fh->height = md.ysize;
  fh->ec_upsampling.assign(nextra, 1);
  fh->ec_blend.assign(nextra, BlendingInfo());
  RestorationFilter& rf = fh->rf;
  memset(&rf, 0, sizeof rf);
  rf.gab = 1;
  rf.epf_iters = 2;
  for (int c = 0; c < 3; ++c) {
    rf.gab_w1[c] = 0.115169525f;
    rf.gab_w2[c] = 0.061248592f;
  }
  for (int i = 0; i < 8; ++i) rf.epf_sharp_lut[i] = (float) i / 7.0f;
  rf.epf_channel_scale[0] = 40.0f;
  rf.epf_channel_scale[1] = 5.0f;
  rf.epf_channel_scale[2] = 3.5f;
  rf.epf_quant_mul = 0.46f;
  rf.epf_pass0_sigma_scale = 0.9f;
  rf.epf_pass2_sigma_scale = 6.5f;
  rf.epf_border_sad_mul = 2.0f / 3.0f;
  rf.epf_sigma_for_modular = 1.0f;
  bool all_default = br.Read(1);
  if (!all_default) {
    fh->frame_type = br.Read(2);
    fh->encoding = br.Read(1);
    fh->flags = br.U64();
    if (!md.xyb_encoded) fh->do_ycbcr = br.Read(1);
    bool use_lf_frame = fh->flags & 0x20;
    if (fh->do_ycbcr && !use_lf_frame)
      for (int i = 0; i < 3; ++i) fh->jpeg_upsampling[i] = br.Read(2);
    if (!use_lf_frame) {
      fh->upsampling = br.U32(1, 0, 2, 0, 4, 0, 8, 0);
      for (uint32_t i = 0; i < nextra; ++i) fh->ec_upsampling[i] = br.U32(1, 0, 2, 0, 4, 0, 8, 0);
    }
    if (fh->encoding == 1) fh->group_size_shift = br.Read(2);
    if (md.xyb_encoded && fh->encoding == 0) {
      fh->x_qm_scale = br.Read(3);
      fh->b_qm_scale = br.Read(3);
    }
    if (fh->frame_type != 2) {
      fh->num_passes = br.U32(1, 0, 2, 0, 3, 0, 4, 3);
      if (fh->num_passes != 1) {
        uint32_t nds = br.U32(0, 0, 1, 0, 2, 0, 3, 1);
        for (uint32_t i = 0; i + 1 < fh->num_passes; ++i) fh->pass_shift[i] = br.Read(2);
        fh->num_ds = nds;
        for (uint32_t i = 0; i < nds; ++i) fh->pass_downsample[i] = br.U32(1, 0, 2, 0, 4, 0, 8, 0);
        for (uint32_t i = 0; i < nds; ++i) fh->pass_last[i] = br.U32(0, 0, 1, 0, 2, 0, 0, 3);
      }
    }
    if (fh->frame_type == 1) {
      fh->lf_level = 1 + br.Read(2);
    } else {
      fh->have_crop = br.Read(1);
      if (fh->have_crop) {
        if (fh->frame_type != 2) {
          fh->x0 = UnpackSigned(br.U32(0, 8, 256, 11, 2304, 14, 18688, 30));
          fh->y0 = UnpackSigned(br.U32(0, 8, 256, 11, 2304, 14, 18688, 30));
        }
        fh->width = br.U32(0, 8, 256, 11, 2304, 14, 18688, 30);
        fh->height = br.U32(0, 8, 256, 11, 2304, 14, 18688, 30);
      }
    }
    bool normal = fh->frame_type == 0 || fh->frame_type == 3;
    bool full = !fh->have_crop || (fh->x0 <= 0 && fh->y0 <= 0 && (int64_t) fh->width + fh->x0 >= (int64_t) md.xsize &&
                                   (int64_t) fh->height + fh->y0 >= (int64_t) md.ysize);
    if (normal) {
      auto read_blend = [&](BlendingInfo* bi) {
        bi->mode = br.U32(0, 0, 1, 0, 2, 0, 3, 2);
        if (nextra > 0 && (bi->mode == 2 || bi->mode == 3)) bi->alpha_channel = br.U32(0, 0, 1, 0, 2, 0, 3, 3);
        if ((nextra > 0 && (bi->mode == 2 || bi->mode == 3)) || bi->mode == 4) bi->clamp = br.Read(1);  // kMul carries it even without extra channels
        if (bi->mode != 0 || !full) bi->source = br.Read(2);
      };
      read_blend(&fh->blend);
      for (uint32_t i = 0; i < nextra; ++i) read_blend(&fh->ec_blend[i]);
      if (md.have_animation) {
        fh->duration = br.U32(0, 0, 1, 0, 0, 8, 0, 32);
        if (md.have_timecodes) fh->timecode = br.Read(32);
      }
      fh->is_last = br.Read(1);
    } else {
      fh->is_last = false;
    }
    if (fh->frame_type != 1 && !fh->is_last) fh->save_as_reference = br.Read(2);
    if (fh->frame_type != 1) {
      bool resets = full && normal && fh->blend.mode == 0;
      bool can_ref = !fh->is_last && (fh->duration == 0 || fh->save_as_reference != 0);
      if (fh->frame_type == 2 || (resets && can_ref)) fh->save_before_ct = br.Read(1);
    }
    uint32_t name_len = br.U32(0, 0, 0, 4, 16, 5, 48, 10);
    for (uint32_t k = 0; k < name_len; ++k) br.Read(8);
    bool rf_default = br.Read(1);
    if (!rf_default) {
      rf.gab = (uint8_t) br.Read(1);
      if (rf.gab) {
        rf.gab_custom = (uint8_t) br.Read(1);
        if (rf.gab_custom)
          for (int c = 0; c < 3; ++c) {
            rf.gab_w1[c] = br.F16();
            rf.gab_w2[c] = br.F16();
          }
      }
      rf.epf_iters = (uint8_t) br.Read(2);
      if (rf.epf_iters) {
        if (fh->encoding == 0 && br.Read(1))
          for (int i = 0; i < 8; ++i) rf.epf_sharp_lut[i] = br.F16();
        if (br.Read(1)) {
          for (int c = 0; c < 3; ++c) rf.epf_channel_scale[c] = br.F16();
          br.Read(32);
        }
        if (br.Read(1)) {
          if (fh->encoding == 0) rf.epf_quant_mul = br.F16();
          rf.epf_pass0_sigma_scale = br.F16();
          rf.epf_pass2_sigma_scale = br.F16();
          rf.epf_border_sad_mul = br.F16();
        }
        if (fh->encoding == 1) rf.epf_sigma_for_modular = br.F16();
      }
      if (br.U64() != 0) JXLB_FAIL(kParseUnsupported, "restoration filter extensions");
    }
    if (br.U64() != 0) JXLB_FAIL(kParseUnsupported, "frame header extensions");
  }
  if (br.Overrun()) JXLB_FAIL(kParseInvalid, "truncated frame header");
  // derived geometry
  uint32_t W = (fh->width + fh->upsampling - 1) / fh->upsampling;
  uint32_t H = (fh->height + fh->upsampling - 1) / fh->upsampling;
  if (fh->lf_level) {
    uint32_t d = 1u << (3 * fh->lf_level);
    W = (W + d - 1) / d;
    H = (H + d - 1) / d;
  }
  fh->coded_w = W;
  fh->coded_h = H;
  if (W == 0 || H == 0) JXLB_FAIL(kParseInvalid, "empty frame");
  // A crop rectangle comes straight from the bitstream (up to 2^30 + 18687 per side).  Frames are bounded like the image
  // itself (DecodeJpegXlOneShot refuses pictures of 2^31 bytes or more, interop/JxlDecoding.cpp:103-109): nothing larger
  // can be handed back, and the group counts below stay far from 32-bit overflow.
  if (fh->width > (1u << 24) || fh->height > (1u << 24) || (uint64_t) fh->width * fh->height >= (1ull << 29))
    JXLB_FAIL(kParseInvalid, "frame size exceeds the decodable image size");
  fh->group_dim = fh->encoding == 0 ? kGroupDim : (128u << fh->group_size_shift);
  const uint32_t gd = fh->group_dim;
  fh->ngx = (W + gd - 1) / gd;
  fh->ngy = (H + gd - 1) / gd;
  fh->nlfx = (W + 8 * gd - 1) / (8 * gd);
  fh->nlfy = (H + 8 * gd - 1) / (8 * gd);
  fh->num_groups = fh->ngx * fh->ngy;
  fh->num_lf_groups = fh->nlfx * fh->nlfy;
  const uint64_t toc64 = (fh->num_groups == 1 && fh->num_passes == 1) ? 1 : 1 + (uint64_t) fh->num_lf_groups + 1 + (uint64_t) fh->num_groups * fh->num_passes;
  // every TOC entry takes at least 10 bits: a table that cannot fit in what is left of the input is refused before any
  // allocation proportional to it
  if (toc64 > (1u << 24) || toc64 * 10 > (uint64_t) cs_len * 8 - std::min<uint64_t>(br.Position(), (uint64_t) cs_len * 8))
    JXLB_FAIL(kParseInvalid, "truncated TOC");
  fh->toc_entries = (uint32_t) toc64;
  // ---- TOC
  const uint32_t n = fh->toc_entries;
  std::vector<uint32_t> perm;
  bool permuted = br.Read(1);
  if (permuted) {
    ScratchLease arena_mem(1 << 20);
    Arena arena;
    arena.Init(arena_mem.data(), 1u << 20);
    uint32_t coff;
    int st = ParseCode<false>(br, 8, true, arena, &coff);
    if (st != kOk) JXLB_FAIL(st == kErrUnsupported ? kParseUnsupported : kParseInvalid, "TOC permutation code");
    CodeView cv;
    cv.Bind(arena.base + coff);
    const uint32_t wn = cv.lz77 ? (1u << kLz77WindowLog) : 1u;
    ScratchLease window((size_t) wn * 4);
    SymbolReader sr;
    sr.Begin(cv, br, window.as<uint32_t>(), wn - 1);
    perm.resize(n);
    std::vector<uint32_t> temp(n);
    st = ReadPermutation(cv, sr, br, n, 0, perm.data(), temp.data());
    if (st != kOk || !sr.FinalStateOk()) JXLB_FAIL(kParseInvalid, "TOC permutation");
  }
  br.AlignToByte();
  std::vector<uint64_t> sizes(n);
  for (uint32_t i = 0; i < n; ++i) sizes[i] = br.U32(0, 10, 1024, 14, 17408, 22, 4211712, 30);
  br.AlignToByte();
  if (br.Overrun()) JXLB_FAIL(kParseInvalid, "truncated TOC");
  uint64_t base = br.Position() / 8;
  std::vector<uint64_t> offs(n);
  uint64_t acc = base;
  for (uint32_t i = 0; i < n; ++i) {
    offs[i] = acc;
    acc += sizes[i];
  }
  fh->end_byte = acc;
  if (acc > cs_len) JXLB_FAIL(kParseInvalid, "frame sections exceed the codestream (truncated input)");
  fh->sec_bit_begin.resize(n);
  fh->sec_bit_end.resize(n);
  for (uint32_t j = 0; j < n; ++j) {
    uint32_t slot = permuted ? perm[j] : j;  // logical section j lives at bitstream slot perm[j]
    fh->sec_bit_begin[j] = offs[slot] * 8;
    fh->sec_bit_end[j] = (offs[slot] + sizes[slot]) * 8;
  }
  return kParseOk;
}

void NaturalCoeffOrder(uint32_t cx, uint32_t cy, std::vector<uint32_t>* out) {
  if (cy > cx) std::swap(cx, cy);
  const uint32_t xs = cx / cy, xsm = xs - 1;
  uint32_t xss = 0;
  while ((1u << xss) < xs) ++xss;
  const uint32_t N = cx * 8;
  out->assign((size_t) 64 * cx * cy, 0);
  uint32_t cur = cx * cy;
  for (uint32_t i = 0; i < N; ++i) {
    for (uint32_t j = 0; j <= i; ++j) {
      uint32_t x = j, y = i - j;
      if (i & 1) std::swap(x, y);
      if (y & xsm) continue;
      y >>= xss;
      uint32_t val;
      if (x < cx && y < cy) val = y * cx + x;
      else val = cur++;
      (*out)[val] = y * N + x;
    }
  }
  for (uint32_t ip = N - 1; ip > 0; --ip) {
    uint32_t i = ip - 1;
    for (uint32_t j = 0; j <= i; ++j) {
      uint32_t x = N - 1 - (i - j), y = N - 1 - j;
      if (i & 1) std::swap(x, y);
      if (y & xsm) continue;
      y >>= xss;
      (*out)[cur++] = y * N + x;
    }
  }
}

int ParseFrameGlobals(const uint8_t* cs, size_t cs_padded, const ImageMetadata& md, const FrameHeader& fh, FrameGlobals* g,
                      std::string* err) {
  if (fh.flags & ~(uint64_t) 0x80) JXLB_FAIL(kParseUnsupported, "noise / patches / splines / LF-frame flags");
  // frames coded at half resolution (libjxl's encoder: distances of about 10 and more) are upsampled 2x with the default
  // kernel (pixel_stages.h: StageUpsample2); 4x / 8x, custom kernels and upsampled extra / modular channels are refused
  if (fh.upsampling != 1 && (fh.upsampling != 2 || fh.encoding != 0 || md.custom_upsampling || fh.have_crop))
    JXLB_FAIL(kParseUnsupported, "upsampled frame");
  // extra channels: at the frame's own resolution (full, or half together with the colour planes)
  for (uint32_t u : fh.ec_upsampling)
    if (u != fh.upsampling) JXLB_FAIL(kParseUnsupported, "upsampled extra channel");
  // progressive passes: the AC coefficients of a group arrive in num_passes sections that add up; its extra channels are
  // split over the passes by their shifts (FrameDev::pass_min_shift / pass_max_shift)
  if (fh.num_passes != 1 && fh.encoding != 0) JXLB_FAIL(kParseUnsupported, "progressive passes of a modular frame");
  if (fh.frame_type != 0) JXLB_FAIL(kParseUnsupported, "non-regular frame type");
  if (fh.do_ycbcr) JXLB_FAIL(kParseUnsupported, "YCbCr frame");
  BitReader br;
  br.Init(cs, cs_padded, fh.sec_bit_begin[0], fh.sec_bit_end[0]);
  // LfChannelDequantization
  if (!br.Read(1))
    for (int c = 0; c < 3; ++c) g->lf_dequant[c] = br.F16() * (1.0f / 128.0f);  // coded unscaled: the defaults are (1/32, 1/4, 1/2) / 128
  ScratchLease arena_mem(8u << 20);
  Arena arena;
  arena.Init(arena_mem.data(), 8u << 20);
  if (fh.encoding == 0) {
    g->global_scale = br.U32(1, 11, 2049, 11, 4097, 12, 8193, 16);
    g->quant_lf = br.U32(16, 0, 1, 5, 1, 8, 1, 16);
    BlockCtxMap& b = g->bctx;
    memset(&b, 0, sizeof b);
    if (br.Read(1)) {
      static const uint8_t kDefault[39] = {0, 1, 2, 2, 3, 3, 4, 5, 6, 6, 6, 6, 6, 7, 8, 9, 9, 10, 11, 12, 13, 14, 14, 14, 14, 14,
                                           7, 8, 9, 9, 10, 11, 12, 13, 14, 14, 14, 14, 14};
      g->bctx_map.assign(kDefault, kDefault + 39);
      b.num_lf_ctx = 1;
      b.num_ctx = 15;
      b.map_size = 39;
    } else {
      uint32_t nlf = 1;
      for (int c = 0; c < 3; ++c) {
        b.num_lf_thr[c] = br.Read(4);
        for (uint32_t i = 0; i < b.num_lf_thr[c]; ++i) b.lf_thr[c][i] = UnpackSigned(br.U32(0, 4, 16, 8, 272, 16, 65808, 32));
        nlf *= b.num_lf_thr[c] + 1;
      }
      b.num_qf_thr = br.Read(4);
      for (uint32_t i = 0; i < b.num_qf_thr; ++i) b.qf_thr[i] = br.U32(0, 2, 4, 3, 12, 5, 44, 8) + 1;
      b.num_lf_ctx = nlf;
      uint32_t size = 3 * kNumOrders * nlf * (b.num_qf_thr + 1);
      if (nlf > 64 || size > 3 * kNumOrders * 64) JXLB_FAIL(kParseInvalid, "block context map too large");
      g->bctx_map.resize(size);
      uint32_t ncl = 0;
      int st = ReadContextMap(br, size, g->bctx_map.data(), &ncl, arena);
      if (st != kOk) JXLB_FAIL(kParseInvalid, "block context map");
      if (ncl > 16) JXLB_FAIL(kParseInvalid, "too many block contexts");
      b.num_ctx = ncl;
      b.map_size = size;
      arena.used = 0;
    }
    // LfChannelCorrelation
    if (!br.Read(1)) {
      g->cfl.colour_factor = br.U32(84, 0, 256, 0, 2, 8, 258, 16);
      g->cfl.base_x = br.F16();
      g->cfl.base_b = br.F16();
      g->cfl.x_factor_lf = br.Read(8);
      g->cfl.b_factor_lf = br.Read(8);
    }
  }
  // GlobalModular: optional tree + code
  g->has_global_tree = br.Read(1);
  if (g->has_global_tree) {
    uint32_t toff, nn, wp, maxp;
    const uint32_t kMaxNodes = 1u << 18;
    int st = DecodeTree(br, arena, kMaxNodes, &toff, &nn, &wp, &maxp);
    if (st != kOk) JXLB_FAIL(st == kErrBadStream ? kParseInvalid : kParseUnsupported, "global MA tree");
    g->tree_blob.assign(arena.base + toff, arena.base + toff + (size_t) nn * sizeof(TreeNode));
    g->tree_nodes = nn;
    g->tree_uses_wp = wp;
    g->tree_max_property = maxp;
    arena.used = 0;
    uint32_t coff;
    st = ParseCode<false>(br, (nn + 1) / 2, true, arena, &coff);
    if (st != kOk) JXLB_FAIL(st == kErrBadStream ? kParseInvalid : kParseUnsupported, "global modular code");
    const CodeHeader* ch = reinterpret_cast<const CodeHeader*>(arena.base + coff);
    g->tree_code.assign(arena.base + coff, arena.base + coff + ch->total_bytes);
    arena.used = 0;
  }
  if (br.Overrun()) JXLB_FAIL(kParseInvalid, "truncated LfGlobal");
  g->global_modular_bit = br.Position();
  memset(&g->global_mh, 0, sizeof g->global_mh);
  {
    const uint32_t nmod = (fh.encoding == 1 ? (md.color.color_space == 1 ? 1u : 3u) : 0u) + (uint32_t) md.extra.size();
    for (const ExtraChannelInfo& ec : md.extra)
      if (ec.dim_shift != 0) JXLB_FAIL(kParseUnsupported, "subsampled extra channel");
    // Multi-section frames: the header is parsed here.  Single-section frames are walked by one device lane from this
    // position on -- unless their extra channels are squeezed, which is handled on the host (below), the lane resuming
    // after the global modular stream (sq_end_bit).
    bool parse_here = nmod > 0 && fh.toc_entries > 1;
    if (nmod > 0 && fh.toc_entries == 1 && fh.encoding == 0) {
      BitReader peek = br;
      ModularHeader mh;
      if (ReadModularHeader(peek, &mh) == kOk && mh.has_squeeze) parse_here = true;
    }
    if (parse_here) {
      int st = ReadModularHeader(br, &g->global_mh);
      if (st == kErrUnsupported) JXLB_FAIL(kParseUnsupported, "weighted-predictor delta palette or too many modular transforms");
      if (st != kOk || br.Overrun()) JXLB_FAIL(kParseInvalid, "global modular header");
      if (!g->global_mh.has_squeeze && fh.toc_entries > 1) {
        // frame-level RCTs / palettes: replay them on the channel list; the palettes' colours (meta channels) are the only
        // channels of a multi-section frame that live in the global stream, and are decoded here
        st = PlanChannels(&g->global_mh, nmod, &g->chplan);
        if (st == kErrUnsupported) JXLB_FAIL(kParseUnsupported, "nested palette transforms");
        if (st != kOk) JXLB_FAIL(kParseInvalid, "modular transform channel range");
        // the global stream of a multi-section frame holds the palettes' colours and -- when the frame is no larger than one
        // group, i.e. a small progressive picture -- the modular channels themselves
        const bool planes_global = fh.coded_w <= fh.group_dim && fh.coded_h <= fh.group_dim && g->chplan.ncoded > 0;
        if (g->chplan.nb_meta || planes_global) {
          ModularContext mc{};
          Arena a2;
          ScratchLease amem(4u << 20);
          a2.Init(amem.data(), 4u << 20);
          if (g->global_mh.use_global_tree) {
            if (!g->has_global_tree) JXLB_FAIL(kParseInvalid, "global tree missing");
            mc.tree = reinterpret_cast<const TreeNode*>(g->tree_blob.data());
            mc.num_nodes = g->tree_nodes;
            mc.uses_wp = g->tree_uses_wp;
            mc.max_property = g->tree_max_property;
            mc.code.Bind(g->tree_code.data());
          } else {
            uint32_t toff, nn, wp, maxp, coff;
            st = DecodeTree(br, a2, 1u << 16, &toff, &nn, &wp, &maxp);
            if (st == kOk) st = ParseCode<false>(br, (nn + 1) / 2, true, a2, &coff);
            if (st != kOk) JXLB_FAIL(st == kErrBadStream ? kParseInvalid : kParseUnsupported, "global modular stream tree");
            mc.tree = reinterpret_cast<const TreeNode*>(a2.base + toff);
            mc.num_nodes = nn;
            mc.uses_wp = wp;
            mc.max_property = maxp;
            mc.code.Bind(a2.base + coff);
          }
          g->meta_data.assign((size_t) g->chplan.meta_ints + 1, 0);
          const uint32_t nglobal = g->chplan.nb_meta + (planes_global ? g->chplan.ncoded : 0);
          std::vector<ModChannel> chs(nglobal);
          uint32_t maxw = planes_global ? fh.coded_w : 0;
          if (planes_global) {
            g->global_planes.assign((size_t) g->chplan.ncoded * fh.coded_w * fh.coded_h, 0);
            for (uint32_t c = 0; c < g->chplan.ncoded; ++c) {
              ModChannel& mcn = chs[g->chplan.nb_meta + c];
              mcn.data = g->global_planes.data() + (size_t) c * fh.coded_w * fh.coded_h;
              mcn.w = fh.coded_w;
              mcn.h = fh.coded_h;
              mcn.stride = fh.coded_w;
            }
          }
          for (uint32_t c = 0; c < g->chplan.nb_meta; ++c) {
            const ModTransform& tr = g->global_mh.tr[g->chplan.meta_tr[c]];
            chs[c].data = g->meta_data.data() + tr.meta_off;
            chs[c].w = PaletteWidth(tr);
            chs[c].h = tr.num_c;
            chs[c].stride = PaletteWidth(tr);
            maxw = std::max(maxw, PaletteWidth(tr));
          }
          std::vector<int32_t> scratch(ModFastScratch::Ints(maxw + 8));
          ScratchLease lz((size_t) 4 << 20);
          st = DecodeModularChannelsFast(br, mc, g->global_mh.wp, chs.data(), nglobal, 0, scratch.data(), lz.as<uint32_t>(), (1u << 20) - 1);
          if (st == kErrUnsupported || st == kErrScratch) JXLB_FAIL(kParseUnsupported, "global modular stream");
          if (st != kOk || br.Overrun()) JXLB_FAIL(kParseInvalid, "global modular stream");
        }
      }
      if (g->global_mh.has_squeeze) {
        // lossy extra channels of a VarDCT frame (squeeze.h); anything wider is refused
        if (fh.encoding != 0 || g->global_mh.nb_transforms != 1) JXLB_FAIL(kParseUnsupported, "squeeze transform outside a VarDCT frame's extra channels");
        std::vector<SqueezeParam> params;
        if (g->global_mh.nb_squeeze == 0) {
          params = DefaultSqueezeParams(nmod, fh.coded_w, fh.coded_h);
        } else {
          for (uint32_t k = 0; k < g->global_mh.nb_squeeze; ++k)
            params.push_back({g->global_mh.sq[k].horizontal != 0, g->global_mh.sq[k].in_place != 0, g->global_mh.sq[k].begin_c, g->global_mh.sq[k].num_c});
        }
        if (!SqueezeLayout(params, nmod, fh.coded_w, fh.coded_h, &g->sq)) JXLB_FAIL(kParseInvalid, "squeeze parameters");
        g->squeeze = true;
        const uint32_t gd = fh.group_dim;
        uint32_t ng = 0;
        while (ng < g->sq.channels.size() && g->sq.channels[ng].w <= gd && g->sq.channels[ng].h <= gd) ++ng;
        g->sq_global = ng;
        uint32_t per_group = 0, per_lf_group = 0;
        for (uint32_t c = ng; c < g->sq.channels.size(); ++c) {
          const SqChannel& sc = g->sq.channels[c];
          if (std::min(sc.hshift, sc.vshift) >= 3) ++per_lf_group;
          else ++per_group;
        }
        if (per_group > 8 || per_lf_group > 8) JXLB_FAIL(kParseUnsupported, "more than 8 squeezed channels per group");
        // the global stream's channels follow the header: decode them here
        ModularContext mc{};
        Arena a2;
        ScratchLease amem(4u << 20);
        a2.Init(amem.data(), 4u << 20);
        if (g->global_mh.use_global_tree) {
          if (!g->has_global_tree) JXLB_FAIL(kParseInvalid, "global tree missing");
          mc.tree = reinterpret_cast<const TreeNode*>(g->tree_blob.data());
          mc.num_nodes = g->tree_nodes;
          mc.uses_wp = g->tree_uses_wp;
          mc.max_property = g->tree_max_property;
          mc.code.Bind(g->tree_code.data());
        } else {
          uint32_t toff, nn, wp, maxp, coff;
          st = DecodeTree(br, a2, 1u << 16, &toff, &nn, &wp, &maxp);
          if (st == kOk) st = ParseCode<false>(br, (nn + 1) / 2, true, a2, &coff);
          if (st != kOk) JXLB_FAIL(st == kErrBadStream ? kParseInvalid : kParseUnsupported, "global modular stream tree");
          mc.tree = reinterpret_cast<const TreeNode*>(a2.base + toff);
          mc.num_nodes = nn;
          mc.uses_wp = wp;
          mc.max_property = maxp;
          mc.code.Bind(a2.base + coff);
        }
        size_t total = 0;
        for (uint32_t c = 0; c < ng; ++c) total += (size_t) g->sq.channels[c].w * g->sq.channels[c].h;
        g->sq_global_data.assign(total + 1, 0);
        std::vector<ModChannel> chs(ng);
        size_t o = 0;
        for (uint32_t c = 0; c < ng; ++c) {
          chs[c].data = g->sq_global_data.data() + o;
          chs[c].w = g->sq.channels[c].w;
          chs[c].h = g->sq.channels[c].h;
          chs[c].stride = chs[c].w;
          o += (size_t) chs[c].w * chs[c].h;
        }
        std::vector<int32_t> scratch(ModFastScratch::Ints(gd + 8));
        ScratchLease lz((size_t) 4 << 20);
        if (ng) {
          st = DecodeModularChannelsFast(br, mc, g->global_mh.wp, chs.data(), ng, 0, scratch.data(), lz.as<uint32_t>(), (1u << 20) - 1);
          if (st == kErrUnsupported || st == kErrScratch) JXLB_FAIL(kParseUnsupported, "global modular stream");
          if (st != kOk || br.Overrun()) JXLB_FAIL(kParseInvalid, "global modular stream");
        }
        g->sq_end_bit = br.Position();
      }
    }
  }
  // HfGlobal for multi-section VarDCT frames
  if (fh.encoding == 0 && fh.toc_entries > 1) {
    uint32_t sec = 1 + fh.num_lf_groups;
    BitReader hbr;
    hbr.Init(cs, cs_padded, fh.sec_bit_begin[sec], fh.sec_bit_end[sec]);
    HfGlobalOut out;
    ScratchLease perm_scratch((size_t) 2 * 65536 * 4);
    int st = ParseHfGlobal(hbr, fh.num_groups, g->bctx.num_ctx, NaturalOrderPoolHost(), arena, perm_scratch.as<uint32_t>(), &out);
    if (st == kErrUnsupported) JXLB_FAIL(kParseUnsupported, "HfGlobal: custom quant tables");
    if (st != kOk || hbr.Overrun()) JXLB_FAIL(kParseInvalid, "HfGlobal");
    g->num_hf_presets = out.num_hf_presets;
    g->used_orders = out.used_orders;
    g->orders = out.orders;
    g->order_pool.assign(reinterpret_cast<const uint16_t*>(arena.base + out.order_pool_off),
                         reinterpret_cast<const uint16_t*>(arena.base + out.order_pool_off) + out.order_pool_entries);
    const CodeHeader* ch = reinterpret_cast<const CodeHeader*>(arena.base + out.ac_code_off);
    g->ac_code.assign(arena.base + out.ac_code_off, arena.base + out.ac_code_off + ch->total_bytes);
    for (uint32_t ps = 1; ps < fh.num_passes; ++ps) {
      HfGlobalOut po;
      po.num_hf_presets = out.num_hf_presets;
      arena.used = 0;  // the previous pass's tables were copied out
      st = ParseHfPass(hbr, g->bctx.num_ctx, NaturalOrderPoolHost(), arena, perm_scratch.as<uint32_t>(), &po);
      if (st != kOk || hbr.Overrun()) JXLB_FAIL(st == kErrUnsupported || st == kErrScratch ? kParseUnsupported : kParseInvalid, "HfGlobal pass");
      FrameGlobals::ExtraPass ep;
      ep.used_orders = po.used_orders;
      ep.orders = po.orders;
      ep.order_pool.assign(reinterpret_cast<const uint16_t*>(arena.base + po.order_pool_off),
                           reinterpret_cast<const uint16_t*>(arena.base + po.order_pool_off) + po.order_pool_entries);
      const CodeHeader* pch = reinterpret_cast<const CodeHeader*>(arena.base + po.ac_code_off);
      ep.ac_code.assign(arena.base + po.ac_code_off, arena.base + po.ac_code_off + pch->total_bytes);
      g->extra_passes.push_back(std::move(ep));
    }
    g->hf_parsed = true;
    g->hf_global_end_bit = hbr.Position();
  }
  return kParseOk;
}

}  // namespace jxlb
