// TEST SUPPORT (host only): a synthetic VarDCT frame that consists of ONE block of a given strategy, so that tests can push
// chosen quantised coefficients and LF samples through the reconstruction kernels (the encoder never emits DCT128 / DCT256,
// so no file reaches those transforms).  Used by jxlb_test_recon_block (c_api.cu) and by tests/hostemu.
#pragma once
#include <cstring>
#include <vector>

#include "frame_parser.h"
#include "plan.h"

namespace jxlb {

inline void MakeSingleBlockPlan(uint32_t strategy, uint32_t global_scale, ImageMetadata* md, FrameHeader* fh, FrameGlobals* g,
                                FramePlan* plan) {
  const uint32_t W = 8 * StrategyCellsX(strategy), H = 8 * StrategyCellsY(strategy);
  md->xsize = W;
  md->ysize = H;
  fh->encoding = 0;
  fh->width = fh->coded_w = W;
  fh->height = fh->coded_h = H;
  fh->group_dim = 256;
  fh->ngx = (W + 255) / 256;
  fh->ngy = (H + 255) / 256;
  fh->nlfx = fh->nlfy = 1;
  fh->num_groups = fh->ngx * fh->ngy;
  fh->num_lf_groups = 1;
  fh->num_passes = 1;
  fh->toc_entries = fh->num_groups == 1 ? 1 : 2 + fh->num_lf_groups + fh->num_groups;
  fh->x_qm_scale = fh->b_qm_scale = 2;  // factor 1
  memset(&fh->rf, 0, sizeof fh->rf);
  fh->sec_bit_begin.assign(fh->toc_entries, 0);
  fh->sec_bit_end.assign(fh->toc_entries, 0);
  g->global_scale = global_scale;
  g->quant_lf = 1;
  g->cfl = CflParams{84, 0.f, 0.f, 128, 128};  // no chroma from luma: every channel is its own dequantised array
  MakeFramePlan(*md, *fh, *g, 16, plan);
}

// f: a FrameDev bound to HOST-addressable regions.  q: [3][8 cy][8 cx] quantised coefficients in plane layout (row =
// vertical frequency); lf: [3][cy][cx] dequantised LF samples.
inline void FillSingleBlock(const FrameDev& f, uint32_t strategy, const int16_t* q, const float* lf, uint32_t hf_mul) {
  const uint32_t cx = StrategyCellsX(strategy), cy = StrategyCellsY(strategy), R = 8 * cy, C = 8 * cx;
  for (uint32_t y = 0; y < cy; ++y)
    for (uint32_t x = 0; x < cx; ++x) {
      const size_t ci = (size_t) y * f.w8 + x;
      f.cell_strategy[ci] = (uint8_t) (strategy | ((x | y) == 0 ? 0x80u : 0u));
      f.cell_hfmul[ci] = (uint16_t) hf_mul;
      f.cell_off[ci] = (uint16_t) ((y << 8) | x);
      f.cell_sharp[ci] = 0;
    }
  const size_t cplane = (size_t) f.coef_h * f.coef_stride, lfplane = (size_t) f.h8 * f.lf_stride;
  for (uint32_t c = 0; c < 3; ++c) {
    for (uint32_t r = 0; r < R; ++r) memcpy(f.coef + c * cplane + (size_t) r * f.coef_stride, q + ((size_t) c * R + r) * C, C * sizeof(int16_t));
    for (uint32_t y = 0; y < cy; ++y) memcpy(f.lf + c * lfplane + (size_t) y * f.lf_stride, lf + ((size_t) c * cy + y) * cx, cx * sizeof(float));
  }
  f.large_list[0] = 0;
  *f.frame_bad = 0;
}

}  // namespace jxlb
