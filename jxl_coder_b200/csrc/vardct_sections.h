// Section decoders of a JPEG XL frame, host + device (see hd.h): HfGlobal parse, LfGroup (LF coefficients + HF
// metadata + block placement), PassGroup (AC coefficients + per-group modular data), global modular data.
// On the GPU each section is one serial stream handled by one lane; single-section frames (<= one group) chain all of
// them in one lane because their sub-streams are concatenated bit-wise.
// Replaces, for this path, libjxl 0.12.0's section decoding behind the reference's DecodeJpegXlOneShot
// (/root/reference/jxlcoder/src/main/cpp/interop/JxlDecoding.cpp:74-175).  Format digest: SURVEY.md App. B.5-B.7.
#pragma once
#include "frame.h"
#include "modular_tight.h"

namespace jxlb {

// Natural coefficient orders for the 13 order ids, one uint16 position per coded coefficient.
struct NaturalOrders {
  const uint16_t* pool;
  uint32_t offset[kNumOrders];
  uint32_t size[kNumOrders];
};
const NaturalOrders& NaturalOrderPoolHost();  // host only; defined in natural_orders.cc

static constexpr uint32_t kOrderInFramePool = 0x80000000u;

struct HfGlobalOut {
  uint32_t num_hf_presets, used_orders;
  OrderTableIndex orders;        // natural-pool offset, or kOrderInFramePool | offset into this frame's pool (uint16 units)
  uint32_t order_pool_off;       // arena offset of the frame's uint16 order pool
  uint32_t order_pool_entries;   // uint16 entries
  uint32_t ac_code_off;          // arena offset of the AC code blob
};

// perm_scratch: 2 * 65536 uint32_t.
// One pass of HfGlobal: the coefficient orders the pass uses and its AC code (out->num_hf_presets must be set).
JXLB_HD_NOINLINE int ParseHfPass(BitReader& br, uint32_t nb_block_ctx, const NaturalOrders& nat, Arena& arena, uint32_t* perm_scratch,
                                 HfGlobalOut* out) {
  uint32_t used = br.U32(0x5F, 0, 0x13, 0, 0, 0, 0, 13);
  out->used_orders = used;
  for (uint32_t o = 0; o < kNumOrders; ++o)
    for (int c = 0; c < 3; ++c) out->orders.offset[o][c] = nat.offset[o];
  out->order_pool_off = 0;
  out->order_pool_entries = 0;
  if (used) {
    uint32_t total = 0;
    for (uint32_t o = 0; o < kNumOrders; ++o)
      if (used >> o & 1) total += 3 * nat.size[o];
    uint32_t poff = arena.Alloc(total * 2, 16);
    if (poff == 0xFFFFFFFFu) return kErrScratch;
    uint16_t* pool = reinterpret_cast<uint16_t*>(arena.base + poff);
    out->order_pool_off = poff;
    out->order_pool_entries = total;
    uint32_t saved = arena.used;
    uint32_t coff;
    int st = ParseCode<false>(br, 8, true, arena, &coff);
    if (st != kOk) return st;
    CodeView cv;
    cv.Bind(arena.base + coff);
    uint32_t* win = nullptr;
    uint32_t wmask = 0;
    if (cv.lz77) {
      uint32_t wo = arena.Alloc(4u << 18, 16);
      if (wo == 0xFFFFFFFFu) return kErrScratch;
      win = reinterpret_cast<uint32_t*>(arena.base + wo);
      wmask = (1u << 18) - 1;
    }
    SymbolReader sr;
    sr.Begin(cv, br, win, wmask);
    uint32_t cur = 0;
    for (uint32_t o = 0; o < kNumOrders; ++o) {
      if (!(used >> o & 1)) continue;
      uint32_t s = OrderRepresentative(o);
      uint32_t llf = StrategyCellsX(s) * StrategyCellsY(s);
      uint32_t size = 64 * llf;
      const uint16_t* natural = nat.pool + nat.offset[o];
      for (int c = 0; c < 3; ++c) {
        st = ReadPermutation(cv, sr, br, size, llf, perm_scratch, perm_scratch + 65536);
        if (st != kOk) return st;
        for (uint32_t k = 0; k < size; ++k) pool[cur + k] = natural[perm_scratch[k]];
        out->orders.offset[o][c] = kOrderInFramePool | cur;
        cur += size;
      }
    }
    if (!sr.FinalStateOk()) return kErrBadStream;
    arena.used = saved;
  }
  uint32_t nctx = kContextsPerBlockCtx * nb_block_ctx * out->num_hf_presets;
  return ParseCode<false>(br, nctx, true, arena, &out->ac_code_off);
}

// HfGlobal of a single-pass frame (quant-table flag, preset count, the one pass).  Progressive frames call ParseHfPass
// once more per further pass (frame_parser.cc).
JXLB_HD_NOINLINE int ParseHfGlobal(BitReader& br, uint32_t num_groups, uint32_t nb_block_ctx, const NaturalOrders& nat,
                                   Arena& arena, uint32_t* perm_scratch, HfGlobalOut* out) {
  if (!br.Read(1)) return kErrUnsupported;  // custom quant tables
  out->num_hf_presets = br.Read(CeilLog2(num_groups)) + 1;
  return ParseHfPass(br, nb_block_ctx, nat, arena, perm_scratch, out);
}

// ---- helpers shared by the section decoders ---------------------------------------------------------------------------
struct StreamScratch {
  Arena arena;             // local trees / codes
  int32_t* wp;             // ModFastScratch::Ints(max channel width) ints
  uint32_t wp_ints = 0;    // capacity of wp (0: not recorded; sized by the caller for the widest group channel)
  uint32_t* lz77;          // optional LZ77 window (power-of-two entries) or nullptr
  uint32_t lz77_mask;
  uint8_t* nzmap;          // 3 * 32 * 32 bytes for the AC non-zero context map
  uint8_t* fast = nullptr;      // optional low-latency memory (shared memory on the device): [code blob | modular scratch]
  uint32_t fast_code_bytes = 0; // capacity of the code-blob part
  uint32_t fast_ints = 0;       // capacity (int32) of the modular-scratch part
};

// Reads a modular sub-stream's GroupHeader and resolves its tree + code (global or local, built in scratch).
JXLB_HD_NOINLINE int BeginModularStream(BitReader& br, const FrameDev& f, StreamScratch& s, uint32_t max_local_nodes,
                                        ModularHeader* mh, ModularContext* mc, bool allow_palette = false) {
  int st = ReadModularHeader(br, mh);
  if (st != kOk) return st;
  if (mh->has_squeeze) return kErrUnsupported;  // squeeze is only handled in a frame's global header (host)
  if (!allow_palette)
    for (uint32_t t = 0; t < mh->nb_transforms; ++t)
      if (mh->tr[t].id == 1) return kErrUnsupported;  // streams decoded straight into fixed planes (LF, HF metadata)
  if (mh->use_global_tree) {
    if (!f.global_tree || !f.global_code) return kErrBadStream;
    mc->tree = f.global_tree;
    mc->num_nodes = f.global_tree_nodes;
    mc->uses_wp = f.global_tree_uses_wp;
    mc->max_property = f.global_tree_max_property;
    mc->code.Bind(f.global_code);
  } else {
    uint32_t toff, nn, wp, maxp;
    st = DecodeTree(br, s.arena, max_local_nodes, &toff, &nn, &wp, &maxp);
    if (st != kOk) return st;
    uint32_t coff;
    st = ParseCode<false>(br, (nn + 1) / 2, true, s.arena, &coff);
    if (st != kOk) return st;
    mc->tree = reinterpret_cast<const TreeNode*>(s.arena.base + toff);
    mc->num_nodes = nn;
    mc->uses_wp = wp;
    mc->max_property = maxp;
    mc->code.Bind(s.arena.base + coff);
  }
  // stage the code (context map + alias tables) in low-latency memory when it fits
  if (s.fast) {
    const CodeHeader* chh = reinterpret_cast<const CodeHeader*>(mc->code.blob);
    if (chh->total_bytes <= s.fast_code_bytes) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(mc->code.blob);
      uint32_t* dst = reinterpret_cast<uint32_t*>(s.fast);
      for (uint32_t i = 0; i < chh->total_bytes / 4; ++i) dst[i] = src[i];
      mc->code.Bind(s.fast);
    }
  }
  return kOk;
}

// Applies the inverse RCTs of a stream header in place (serial; used for the small LF / metadata planes and by the
// per-group modular data, whose group-local RCT must be undone before the frame-global transforms).
JXLB_HD void ApplyInverseRcts(const ModularHeader& mh, ModChannel* ch, uint32_t nch) {
  for (int t = (int) mh.nb_transforms - 1; t >= 0; --t) {
    const ModTransform& tr = mh.tr[t];
    if (tr.id != 0 || tr.begin_c + 3 > nch) continue;
    ModChannel& a = ch[tr.begin_c];
    ModChannel& b = ch[tr.begin_c + 1];
    ModChannel& c = ch[tr.begin_c + 2];
    uint32_t perm[3];
    RctPermutation(tr.rct_type, perm);
    ModChannel* dst[3] = {&ch[tr.begin_c + perm[0]], &ch[tr.begin_c + perm[1]], &ch[tr.begin_c + perm[2]]};
    for (uint32_t y = 0; y < a.h; ++y) {
      for (uint32_t x = 0; x < a.w; ++x) {
        int32_t v0 = a.data[(size_t) y * a.stride + x];
        int32_t v1 = b.data[(size_t) y * b.stride + x];
        int32_t v2 = c.data[(size_t) y * c.stride + x];
        InverseRctPixel(tr.rct_type, v0, v1, v2);
        dst[0]->data[(size_t) y * dst[0]->stride + x] = v0;
        dst[1]->data[(size_t) y * dst[1]->stride + x] = v1;
        dst[2]->data[(size_t) y * dst[2]->stride + x] = v2;
      }
    }
  }
}

// Decodes a stream whose result is `nplanes` equally sized planes, honouring the transforms of its header: palette
// colours (meta channels, first in the stream) go to scratch, the remaining coded channels straight into the planes the
// inverse transforms then expand in place (PlanChannels, modular.h).
JXLB_HD_NOINLINE int DecodeTransformedStream(BitReader& br, StreamScratch& s, ModularHeader& mh, const ModularContext& mc,
                                             const ModChannel* planes, uint32_t nplanes, uint32_t stream_id, uint32_t bit_depth) {
  ChannelPlan cp;
  int st = PlanChannels(&mh, nplanes, &cp);
  if (st != kOk) return st;
  int32_t* meta = nullptr;
  if (cp.meta_ints) {
    const uint32_t mo = s.arena.Alloc(cp.meta_ints * 4u, 16);
    if (mo == 0xFFFFFFFFu) return kErrScratch;
    meta = reinterpret_cast<int32_t*>(s.arena.base + mo);
  }
  ModChannel list[kMaxModPlanes + kMaxTransforms];
  uint32_t n = 0, maxw = 0;
  for (uint32_t i = 0; i < cp.nb_meta; ++i) {
    const ModTransform& tr = mh.tr[cp.meta_tr[i]];
    list[n].data = meta + tr.meta_off;
    list[n].w = PaletteWidth(tr);
    list[n].h = tr.num_c;
    list[n].stride = PaletteWidth(tr);
    if (PaletteWidth(tr) > maxw) maxw = PaletteWidth(tr);
    ++n;
  }
  for (uint32_t i = 0; i < cp.ncoded; ++i) {
    list[n] = planes[cp.coded_plane[i]];
    if (list[n].w > maxw) maxw = list[n].w;
    ++n;
  }
  // the caller's predictor scratch is sized for the planes; a palette wider than that borrows from the arena
  int32_t* wp = s.wp;
  if (cp.nb_meta && (s.wp_ints == 0 || ModFastScratch::Ints(maxw) > s.wp_ints)) {
    const uint32_t wo = s.arena.Alloc(ModFastScratch::Ints(maxw) * 4u, 16);
    if (wo == 0xFFFFFFFFu) return kErrScratch;
    wp = reinterpret_cast<int32_t*>(s.arena.base + wo);
  }
  st = DecodeModularChannelsFast(br, mc, mh.wp, list, n, stream_id, wp, s.lz77, s.lz77_mask,
                                 s.fast ? reinterpret_cast<int32_t*>(s.fast + s.fast_code_bytes) : nullptr, s.fast_ints);
  if (st != kOk) return st;
  return ApplyInverseTransforms(mh.tr, mh.nb_transforms, planes, meta, bit_depth);
}

JXLB_HD void LfGroupRect(const FrameDev& f, uint32_t lfg, uint32_t* cx0, uint32_t* cy0, uint32_t* w8, uint32_t* h8,
                         uint32_t* w64, uint32_t* h64) {
  uint32_t gx = lfg % f.nlfx, gy = lfg / f.nlfx;
  *cx0 = gx * kLfGroupCells;
  *cy0 = gy * kLfGroupCells;
  uint32_t px0 = gx * kLfGroupCells * 8, py0 = gy * kLfGroupCells * 8;
  uint32_t pw = f.width - px0 < kLfGroupCells * 8 ? f.width - px0 : kLfGroupCells * 8;
  uint32_t ph = f.height - py0 < kLfGroupCells * 8 ? f.height - py0 : kLfGroupCells * 8;
  *w8 = (pw + 7) / 8;
  *h8 = (ph + 7) / 8;
  *w64 = (pw + 63) / 64;
  *h64 = (ph + 63) / 64;
}

// Squeezed extra channels (squeeze.h) that a modular sub-stream covering the pixel rectangle (x0, y0, dim x dim) codes:
// the pyramid channels outside the global stream with min_shift <= min(hshift, vshift) <= max_shift, each restricted to
// the rectangle scaled by its shifts.  LF groups carry shifts >= 3, pass groups (single pass) shifts 0..2.
JXLB_HD int CollectSqueezeChannels(const FrameDev& f, uint32_t x0, uint32_t y0, uint32_t dim, uint32_t min_shift, uint32_t max_shift,
                                   ModChannel* ch, uint32_t* nch_out) {
  uint32_t nch = 0;
  for (uint32_t c = f.sq_global; c < f.sq_nch; ++c) {
    const SqChannel sc = f.sq_ch[c];
    const uint32_t shift = sc.hshift < sc.vshift ? sc.hshift : sc.vshift;
    if (shift < min_shift || shift > max_shift) continue;
    const uint32_t rx = x0 >> sc.hshift, ry = y0 >> sc.vshift;
    if (rx >= sc.w || ry >= sc.h) continue;
    uint32_t rw = dim >> sc.hshift, rh = dim >> sc.vshift;
    if (rw > sc.w - rx) rw = sc.w - rx;
    if (rh > sc.h - ry) rh = sc.h - ry;
    if (!rw || !rh) continue;
    if (nch == 8) return kErrUnsupported;
    ch[nch].data = f.sq_buf + sc.off + (size_t) ry * sc.w + rx;
    ch[nch].w = rw;
    ch[nch].h = rh;
    ch[nch].stride = sc.w;
    ++nch;
  }
  *nch_out = nch;
  return kOk;
}

// LfGroup section of a VarDCT frame (App. B.7): LF coefficients, HF metadata, block placement.
// place_blocks = false: stop after the entropy-coded data (the CUDA kernel then places the blocks with the whole warp,
// PlaceBlocksWarp in kernels.cu, which must produce exactly what the serial loop below produces).
JXLB_HD_NOINLINE int DecodeLfGroupSection(BitReader& br, const FrameDev& f, uint32_t lfg, StreamScratch& s,
                                          uint32_t max_local_nodes, bool place_blocks = true) {
  uint32_t cx0, cy0, w8, h8, w64, h64;
  LfGroupRect(f, lfg, &cx0, &cy0, &w8, &h8, &w64, &h64);
  const uint32_t nlf = f.num_lf_groups;
  f.lf_extra_precision[lfg] = br.Read(2);
  // ---- LF coefficients: 3 channels, coded Y, X, B
  ModularHeader mh;
  ModularContext mc;
  uint32_t arena_mark = s.arena.used;
  int st = BeginModularStream(br, f, s, max_local_nodes, &mh, &mc);
  if (st != kOk) return st;
  ModChannel ch[4];
  for (int c = 0; c < 3; ++c) {
    ch[c].data = f.lf_quant + (size_t) c * f.h8 * f.lf_stride + (size_t) cy0 * f.lf_stride + cx0;
    ch[c].w = w8;
    ch[c].h = h8;
    ch[c].stride = f.lf_stride;
  }
  st = DecodeModularChannelsTight(br, mc, mh.wp, ch, 3, 1 + lfg, s.wp, s.lz77, s.lz77_mask,
                                 s.fast ? reinterpret_cast<int32_t*>(s.fast + s.fast_code_bytes) : nullptr, s.fast_ints);
  if (st != kOk) return st;
  ApplyInverseRcts(mh, ch, 3);
  s.arena.used = arena_mark;
  // ---- ModularLfGroup: squeezed extra channels with both shifts >= 3 (images wider or taller than 2048 pixels)
  if (f.sq_nch) {
    ModChannel sch[8];
    uint32_t nsq = 0;
    st = CollectSqueezeChannels(f, cx0 * 8, cy0 * 8, kLfGroupCells * 8, 3, 1000, sch, &nsq);
    if (st != kOk) return st;
    if (nsq) {
      st = BeginModularStream(br, f, s, max_local_nodes, &mh, &mc);
      if (st != kOk) return st;
      if (mh.nb_transforms) return kErrUnsupported;
      st = DecodeModularChannelsFast(br, mc, mh.wp, sch, nsq, 1 + nlf + lfg, s.wp, s.lz77, s.lz77_mask,
                                     s.fast ? reinterpret_cast<int32_t*>(s.fast + s.fast_code_bytes) : nullptr, s.fast_ints);
      if (st != kOk) return st;
      s.arena.used = arena_mark;
    }
  }
  // ---- HF metadata
  uint32_t nb_blocks = br.Read(CeilLog2(w8 * h8)) + 1;
  if (nb_blocks > w8 * h8) return kErrBadStream;
  f.nb_blocks[lfg] = nb_blocks;
  st = BeginModularStream(br, f, s, max_local_nodes, &mh, &mc);
  if (st != kOk) return st;
  uint32_t tx0 = cx0 / 8, ty0 = cy0 / 8;
  ch[0].data = f.xfromy + (size_t) ty0 * f.w64 + tx0;
  ch[0].w = w64;
  ch[0].h = h64;
  ch[0].stride = f.w64;
  ch[1] = ch[0];
  ch[1].data = f.bfromy + (size_t) ty0 * f.w64 + tx0;
  int32_t* binfo = f.blockinfo + f.blockinfo_off[lfg];
  ch[2].data = binfo;
  ch[2].w = nb_blocks;
  ch[2].h = 2;
  ch[2].stride = nb_blocks;
  ch[3].data = f.sharpness_i32 + (size_t) cy0 * f.lf_stride + cx0;
  ch[3].w = w8;
  ch[3].h = h8;
  ch[3].stride = f.lf_stride;
  st = DecodeModularChannelsTight(br, mc, mh.wp, ch, 4, 1 + 2 * nlf + lfg, s.wp, s.lz77, s.lz77_mask,
                                 s.fast ? reinterpret_cast<int32_t*>(s.fast + s.fast_code_bytes) : nullptr, s.fast_ints);
  if (st != kOk) return st;
  if (mh.nb_transforms) return kErrUnsupported;
  s.arena.used = arena_mark;
  if (!place_blocks) return br.Overrun() ? kErrTruncated : kOk;
  // ---- block placement: next BlockInfo entry goes to the first uncovered cell in raster order
  for (uint32_t y = 0; y < h8; ++y) {
    uint8_t* row = f.cell_strategy + (size_t) (cy0 + y) * f.w8 + cx0;
    for (uint32_t x = 0; x < w8; ++x) row[x] = 0xFF;
  }
  uint32_t k = 0;
  for (uint32_t y = 0; y < h8; ++y) {
    for (uint32_t x = 0; x < w8; ++x) {
      uint8_t* cell = f.cell_strategy + (size_t) (cy0 + y) * f.w8 + cx0 + x;
      int32_t sh = f.sharpness_i32[(size_t) (cy0 + y) * f.lf_stride + cx0 + x];
      if (sh < 0 || sh > 7) return kErrBadStream;
      f.cell_sharp[(size_t) (cy0 + y) * f.w8 + cx0 + x] = (uint8_t) sh;
      if (*cell != 0xFF) continue;
      if (k >= nb_blocks) return kErrBadStream;
      int32_t t = binfo[k];
      int32_t q = binfo[nb_blocks + k];
      q = 1 + (q < 0 ? 0 : q > 255 ? 255 : q);  // hf_mul is clamped to [1, 256]
      ++k;
      if (t < 0 || t >= kNumStrategies) return kErrBadStream;
      uint32_t bx = StrategyCellsX((uint32_t) t), by = StrategyCellsY((uint32_t) t);
      if (x + bx > w8 || y + by > h8) return kErrBadStream;
      // a varblock may not straddle a 256x256 group boundary
      if ((x % kGroupCells) + bx > kGroupCells || (y % kGroupCells) + by > kGroupCells) return kErrBadStream;
      for (uint32_t yy = 0; yy < by; ++yy) {
        uint8_t* r = f.cell_strategy + (size_t) (cy0 + y + yy) * f.w8 + cx0 + x;
        uint16_t* rq = f.cell_hfmul + (size_t) (cy0 + y + yy) * f.w8 + cx0 + x;
        uint16_t* ro = f.cell_off + (size_t) (cy0 + y + yy) * f.w8 + cx0 + x;
        for (uint32_t xx = 0; xx < bx; ++xx) {
          if (r[xx] != 0xFF) return kErrBadStream;
          r[xx] = (uint8_t) t;
          rq[xx] = (uint16_t) q;
          ro[xx] = (uint16_t) ((yy << 8) | xx);
        }
      }
      *cell = (uint8_t) (t | 0x80);
    }
  }
  if (k != nb_blocks) return kErrBadStream;
  if (br.Overrun()) return kErrTruncated;
  return kOk;
}

// Block context for (order id, hf_mul, channel) -- libjxl's BlockCtxMap::Context; c: 0=X 1=Y 2=B.
JXLB_HD uint32_t BlockContext(const FrameDev& f, uint32_t order, uint32_t hf_mul, uint32_t c, uint32_t lf_idx) {
  uint32_t qi = 0;
  for (uint32_t i = 0; i < f.bctx.num_qf_thr; ++i) qi += hf_mul > f.bctx.qf_thr[i] ? 1 : 0;
  uint32_t idx = c < 2 ? (c ^ 1) : 2;
  idx = idx * kNumOrders + order;
  idx = idx * (f.bctx.num_qf_thr + 1) + qi;
  idx = idx * f.bctx.num_lf_ctx + lf_idx;
  return f.bctx_map[idx];
}

// AC part of a PassGroup section (App. B.7).  Writes the non-zero quantised coefficients of group g into the frame's
// coefficient planes (pre-zeroed), each block's coefficient array occupying the block's pixel rectangle, transposed
// for tall blocks (DESIGN.md "coefficient planes").
// pass: which progressive pass this section is (0 for ordinary frames); the tables of passes > 0 come from f.pass_table and
// their coefficients, coded >> shift, are added to what the earlier passes left in the planes.
JXLB_HD_NOINLINE int DecodeAcGroup(BitReader& br_io, const FrameDev& f, uint32_t g, const NaturalOrders& nat, StreamScratch& s,
                                   uint32_t pass = 0) {
  BitReader br = br_io;  // register copy: the coefficient stores below must not force reloads of the reader state
  const uint8_t* ac_code = f.ac_code;
  const uint16_t* order_pool = f.order_pool;
  const OrderTableIndex* orders = &f.orders;
  uint32_t shift = f.pass_shift0;
  if (pass > 0) {
    const PassDev& pd = f.pass_table[pass - 1];
    ac_code = reinterpret_cast<const uint8_t*>(f.pass_table) + pd.ac_code_rel;
    order_pool = reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(f.pass_table) + pd.order_pool_rel);
    orders = &pd.orders;
    shift = pd.shift;
  }
  const bool accumulate = f.num_passes > 1;
  const uint32_t gx = g % f.ngx, gy = g / f.ngx;
  const uint32_t bx0 = gx * kGroupCells, by0 = gy * kGroupCells;
  const uint32_t bw = f.w8 - bx0 < kGroupCells ? f.w8 - bx0 : kGroupCells;
  const uint32_t bh = f.h8 - by0 < kGroupCells ? f.h8 - by0 : kGroupCells;
  const uint32_t nbc = f.bctx.num_ctx;
  uint32_t hfp = br.Read(CeilLog2(f.num_hf_presets));
  if (hfp >= f.num_hf_presets) return kErrBadStream;
  const uint32_t ctx_off = hfp * kContextsPerBlockCtx * nbc;
  CodeView code;
  code.Bind(ac_code);
  if (code.lz77 && s.lz77 == nullptr) return kErrUnsupported;
  SymbolReader sr;
  sr.Begin(code, br, s.lz77, s.lz77_mask);
  uint8_t* nzmap = s.nzmap;  // [3][32][32]
  const bool has_lf_thr = f.bctx.num_lf_ctx > 1;
  for (uint32_t by = 0; by < bh; ++by) {
    const uint8_t* strat_row = f.cell_strategy + (size_t) (by0 + by) * f.w8 + bx0;
    const uint16_t* mul_row = f.cell_hfmul + (size_t) (by0 + by) * f.w8 + bx0;
    for (uint32_t bx = 0; bx < bw; ++bx) {
      uint32_t sv = strat_row[bx];
      if (!(sv & 0x80) || sv == 0xFF) continue;
      const uint32_t t = sv & 0x7F;
      const uint32_t q = mul_row[bx];
      const uint32_t cx = StrategyCellsX(t), cy = StrategyCellsY(t);
      const uint32_t covered = cx * cy;
      const uint32_t l2 = (uint32_t) FloorLog2(covered);
      const uint32_t size = 64 * covered;
      const uint32_t ord = StrategyOrder(t);
      const uint32_t kc_log = 3 + (uint32_t) FloorLog2(cx > cy ? cx : cy);  // log2 of the coefficient array's columns
      // Square and tall blocks keep their coefficient array transposed in the plane, so that for every block the
      // plane holds F[v][u] (v = vertical, u = horizontal frequency) at (row v, col u) of the block's rectangle.
      const bool tall = cy >= cx;
      uint32_t lf_idx = 0;
      if (has_lf_thr) {
        // bucket = (bx * (nB+1) + bb) * (nY+1) + by over the quantised LF values (planes stored Y, X, B)
        size_t cell = (size_t) (by0 + by) * f.lf_stride + bx0 + bx;
        size_t plane = (size_t) f.h8 * f.lf_stride;
        int32_t vy = f.lf_quant[cell], vx = f.lf_quant[plane + cell], vb = f.lf_quant[2 * plane + cell];
        uint32_t ix = 0, iy = 0, ib = 0;
        for (uint32_t i = 0; i < f.bctx.num_lf_thr[0]; ++i) ix += vx > f.bctx.lf_thr[0][i] ? 1 : 0;
        for (uint32_t i = 0; i < f.bctx.num_lf_thr[1]; ++i) iy += vy > f.bctx.lf_thr[1][i] ? 1 : 0;
        for (uint32_t i = 0; i < f.bctx.num_lf_thr[2]; ++i) ib += vb > f.bctx.lf_thr[2][i] ? 1 : 0;
        lf_idx = (ix * (f.bctx.num_lf_thr[2] + 1) + ib) * (f.bctx.num_lf_thr[1] + 1) + iy;
      }
      for (uint32_t ci = 0; ci < 3; ++ci) {
        const uint32_t c = ci == 0 ? 1 : ci == 1 ? 0 : 2;  // coded Y, X, B
        uint8_t* nzm = nzmap + c * 1024;
        uint32_t pred;
        if (bx == 0 && by == 0) pred = 32;
        else if (bx == 0) pred = nzm[(by - 1) * 32 + bx];
        else if (by == 0) pred = nzm[by * 32 + bx - 1];
        else pred = (nzm[(by - 1) * 32 + bx] + nzm[by * 32 + bx - 1] + 1u) >> 1;
        const uint32_t bc = BlockContext(f, ord, q, c, lf_idx);
        const uint32_t nzc = pred < 8 ? pred : (pred >= 64 ? 36 : 4 + pred / 2);
        uint32_t nz = ReadHybridUint(code, sr, br, ctx_off + nzc * nbc + bc);
        if (nz > size - covered) return kErrBadStream;
        const uint8_t nzv = (uint8_t) ((nz + covered - 1) >> l2);
        for (uint32_t yy = 0; yy < cy; ++yy)
          for (uint32_t xx = 0; xx < cx; ++xx) nzm[(by + yy) * 32 + bx + xx] = nzv;
        if (nz == 0) continue;
        const uint32_t ooff = orders->offset[ord][c];
        const uint16_t* order = (ooff & kOrderInFramePool) ? order_pool + (ooff & ~kOrderInFramePool)
                                                            : nat.pool + ooff;
        int16_t* plane = f.coef + (size_t) c * f.coef_h * f.coef_stride + (size_t) (by0 + by) * 8 * f.coef_stride + (bx0 + bx) * 8;
        const uint32_t h0 = ctx_off + nbc * kNonZeroBuckets + kZeroDensityContexts * bc;
        uint32_t prev = nz > size / 16 ? 0 : 1;
        for (uint32_t k = covered; k < size && nz != 0; ++k) {
          const uint32_t nl = (nz + covered - 1) >> l2;
          const uint32_t ctx = h0 + (ZeroDensityNnzCtx(nl) + ZeroDensityFreqCtx(k >> l2)) * 2 + prev;
          const uint32_t u = ReadHybridUint(code, sr, br, ctx);
          prev = u != 0 ? 1 : 0;
          nz -= prev;
          if (u) {
            int32_t v = UnpackSigned(u);
            if (v > 32767 || v < -32768) return kErrUnsupported;
            const uint32_t pos = order[k];
            uint32_t r = pos >> kc_log, col = pos & ((1u << kc_log) - 1);
            if (tall) {
              uint32_t tmp = r;
              r = col;
              col = tmp;
            }
            if (accumulate) {
              v = plane[(size_t) r * f.coef_stride + col] + v * (1 << shift);
              if (v > 32767 || v < -32768) return kErrUnsupported;
            }
            plane[(size_t) r * f.coef_stride + col] = (int16_t) v;
          }
        }
        if (nz != 0) return kErrBadStream;
      }
    }
  }
  if (!sr.FinalStateOk()) return kErrBadStream;
  if (br.Overrun()) return kErrTruncated;
  br_io = br;
  return kOk;
}

// Per-group modular data of group g (the tail of a PassGroup section, or the whole section of a modular frame):
// all channels of the frame's modular image that were not decoded globally, restricted to the group's rectangle.
JXLB_HD_NOINLINE int DecodeModularGroup(BitReader& br, const FrameDev& f, uint32_t g, StreamScratch& s, uint32_t max_local_nodes,
                                        uint32_t pass = 0) {
  const uint32_t min_shift = f.pass_min_shift[pass], max_shift = f.pass_max_shift[pass];
  if (min_shift == 255) return kOk;  // this pass carries no modular channel
  const uint32_t gd = f.group_dim;
  const uint32_t gx = g % f.ngx, gy = g / f.ngx;
  const uint32_t x0 = gx * gd, y0 = gy * gd;
  ModChannel ch[8];
  uint32_t nch = 0;
  if (f.sq_nch) {
    // squeezed extra channels (squeeze.h); a group that no channel reaches codes nothing, not even a header
    const int cst = CollectSqueezeChannels(f, x0, y0, gd, min_shift, max_shift, ch, &nch);
    if (cst != kOk) return cst;
    if (!nch) return kOk;
  } else {
    // the coded channels left once the frame-level palettes are applied (all of them, without palettes)
    if (f.global_mod_decoded >= f.num_mod_channels) return kOk;
    if (min_shift > 0) return kOk;  // full-resolution channels arrive with the pass that reaches shift 0
    nch = f.num_coded;
    if (nch > 8) return kErrUnsupported;
    const uint32_t w = f.width - x0 < gd ? f.width - x0 : gd;
    const uint32_t h = f.height - y0 < gd ? f.height - y0 : gd;
    for (uint32_t i = 0; i < nch; ++i) {
      ch[i].data = f.mod + (size_t) f.coded_plane[i] * f.height * f.mod_stride + (size_t) y0 * f.mod_stride + x0;
      ch[i].w = w;
      ch[i].h = h;
      ch[i].stride = f.mod_stride;
    }
  }
  ModularHeader mh;
  ModularContext mc;
  uint32_t arena_mark = s.arena.used;
  int st = BeginModularStream(br, f, s, max_local_nodes, &mh, &mc, /*allow_palette=*/!f.sq_nch);
  if (st != kOk) return st;
  const uint32_t stream_id = 1 + 3 * f.num_lf_groups + 17 + f.num_groups * pass + g;
  if (f.sq_nch) {
    if (mh.nb_transforms) return kErrUnsupported;
    st = DecodeModularChannelsFast(br, mc, mh.wp, ch, nch, stream_id, s.wp, s.lz77, s.lz77_mask,
                                   s.fast ? reinterpret_cast<int32_t*>(s.fast + s.fast_code_bytes) : nullptr, s.fast_ints);
  } else {
    st = DecodeTransformedStream(br, s, mh, mc, ch, nch, stream_id, f.bit_depth);  // group-local RCTs / palettes undone here
  }
  if (st != kOk) return st;
  s.arena.used = arena_mark;
  if (br.Overrun()) return kErrTruncated;
  return kOk;
}

// Undoes the squeeze of a frame's extra channels once every group has been decoded (serial version for the CPU
// emulation; the device runs one kernel per step, kernels_post.cu).
inline void UnsqueezeAllSerial(const FrameDev& f) {
  size_t o = 0;
  for (uint32_t c = 0; c < f.sq_global; ++c) {
    const SqChannel sc = f.sq_ch[c];
    for (size_t i = 0; i < (size_t) sc.w * sc.h; ++i) f.sq_buf[sc.off + i] = f.sq_global_data[o + i];
    o += (size_t) sc.w * sc.h;
  }
  for (uint32_t k = 0; k < f.sq_nsteps; ++k) {
    const SqStep st = f.sq_steps[k];
    int32_t* out = st.out_off == 0xFFFFFFFFu ? f.mod + (size_t) st.final_channel * f.height * f.mod_stride : f.sq_buf + st.out_off;
    const uint32_t ow = st.horizontal ? st.avg_w + st.res_w : st.avg_w;
    const uint32_t ostride = st.out_off == 0xFFFFFFFFu ? f.mod_stride : ow;
    if (st.horizontal) {
      for (uint32_t y = 0; y < st.avg_h; ++y)
        InvSqueezeRow(f.sq_buf + st.avg_off + (size_t) y * st.avg_w, f.sq_buf + st.res_off + (size_t) y * st.res_w, out + (size_t) y * ostride,
                      st.avg_w, st.res_w);
    } else {
      for (uint32_t x = 0; x < st.avg_w; ++x)
        InvSqueezeColumn(f.sq_buf + st.avg_off + x, st.avg_w, f.sq_buf + st.res_off + x, st.res_w, out + x, ostride, st.avg_h, st.res_h);
    }
  }
}

// Channels of the frame's modular image decoded inside the global stream (single-section frames: all of them).
JXLB_HD_NOINLINE int DecodeGlobalModular(BitReader& br, const FrameDev& f, StreamScratch& s, uint32_t max_local_nodes) {
  const uint32_t nch = f.global_mod_decoded;
  if (f.num_mod_channels == 0) return kOk;  // no GroupHeader is coded for an empty modular image
  if (nch > 8) return kErrUnsupported;
  ModChannel ch[8];
  for (uint32_t i = 0; i < nch; ++i) {
    ch[i].data = f.mod + (size_t) i * f.height * f.mod_stride;
    ch[i].w = f.width;
    ch[i].h = f.height;
    ch[i].stride = f.mod_stride;
  }
  ModularHeader mh;
  ModularContext mc;
  uint32_t arena_mark = s.arena.used;
  int st = BeginModularStream(br, f, s, max_local_nodes, &mh, &mc, /*allow_palette=*/true);
  if (st != kOk) return st;
  st = DecodeTransformedStream(br, s, mh, mc, ch, nch, 0, f.bit_depth);
  if (st != kOk) return st;
  s.arena.used = arena_mark;
  return kOk;
}

// A frame whose TOC has a single entry: LfGlobal | LfGroup | HfGlobal | PassGroup are concatenated bit-wise, so one
// lane walks them in order.  `f_in` has the LfGlobal fields filled by the host; br starts at f_in.global_modular_bit.
// hf_arena receives the HfGlobal tables (order pool + AC code); perm_scratch: 2 * 65536 uint32_t.
JXLB_HD_NOINLINE int DecodeSingleSectionFrame(const FrameDev& f_in, const NaturalOrders& nat, StreamScratch& s, Arena& hf_arena,
                                              uint32_t* perm_scratch, uint32_t max_local_nodes) {
  FrameDev f = f_in;
  BitReader br;
  int st = kOk;
  if (f.sq_nch) {  // squeezed extra channels: the host decoded the global modular stream (frame_parser.cc)
    br.Init(f.cs, f.cs_bytes, f.sq_end_bit, f.sec_bit_end[0]);
  } else {
    br.Init(f.cs, f.cs_bytes, f.global_modular_bit, f.sec_bit_end[0]);
    st = DecodeGlobalModular(br, f, s, max_local_nodes);
    if (st != kOk) return st;
  }
  if (f.encoding == 0) {
    st = DecodeLfGroupSection(br, f, 0, s, max_local_nodes);
    if (st != kOk) return st;
    HfGlobalOut hf;
    st = ParseHfGlobal(br, f.num_groups, f.bctx.num_ctx, nat, hf_arena, perm_scratch, &hf);
    if (st != kOk) return st;
    f.num_hf_presets = hf.num_hf_presets;
    f.used_orders = hf.used_orders;
    f.orders = hf.orders;
    f.order_pool = reinterpret_cast<const uint16_t*>(hf_arena.base + hf.order_pool_off);
    f.ac_code = hf_arena.base + hf.ac_code_off;
    st = DecodeAcGroup(br, f, 0, nat, s);
    if (st != kOk) return st;
  }
  st = DecodeModularGroup(br, f, 0, s, max_local_nodes);
  if (st != kOk) return st;
  if (br.Overrun()) return kErrTruncated;
  return kOk;
}

}  // namespace jxlb
