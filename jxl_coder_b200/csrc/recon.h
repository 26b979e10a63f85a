// CTA-level reconstruction procedures, host + device: LF dequantisation + adaptive smoothing, and the per-region
// dequant -> chroma-from-luma -> LLF synthesis -> inverse VarDCT of a 64x64-pixel region held in shared memory.
// Every procedure is written as "all threads of a CTA call it with (tid, nthreads)" and separates its phases with a
// sync functor (__syncthreads on the device, a no-op for the single-threaded CPU emulation in tests/hostemu), so the
// exact kernel logic runs on a box without a GPU.
// Replaces, for this path, libjxl 0.12.0 behind the reference's DecodeJpegXlOneShot
// (/root/reference/jxlcoder/src/main/cpp/interop/JxlDecoding.cpp:74-175).  Formulas: SURVEY.md App. B.7.
#pragma once
#include "numeric.h"

namespace jxlb {

// ---- LF ----------------------------------------------------------------------------------------------------------------
struct LfMul {
  float mul[3];   // X, Y, B: m_lf[c] * (65536 / global_scale) / quant_lf
  float kx, kb;   // chroma-from-luma factors for LF
};
JXLB_HD LfMul MakeLfMul(const FrameDev& f) {
  LfMul m;
  const float inv_gs = 65536.0f / (float) f.global_scale;
  for (int c = 0; c < 3; ++c) m.mul[c] = f.lf_dequant[c] * inv_gs / (float) f.quant_lf;
  m.kx = f.cfl.base_x + ((float) f.cfl.x_factor_lf - 128.0f) / (float) f.cfl.colour_factor;
  m.kb = f.cfl.base_b + ((float) f.cfl.b_factor_lf - 128.0f) / (float) f.cfl.colour_factor;
  return m;
}

// Dequantised LF (X, Y, B) of cell (cx, cy) before smoothing.
JXLB_HD void LfDequantCell(const FrameDev& f, const LfMul& m, uint32_t cx, uint32_t cy, float out[3]) {
  const uint32_t lfg = (cy / kLfGroupCells) * f.nlfx + cx / kLfGroupCells;
  const float ep = 1.0f / (float) (1u << f.lf_extra_precision[lfg]);
  const size_t plane = (size_t) f.h8 * f.lf_stride, i = (size_t) cy * f.lf_stride + cx;
  const float y = (float) f.lf_quant[i] * (m.mul[1] * ep);
  const float x = (float) f.lf_quant[plane + i] * (m.mul[0] * ep);
  const float b = (float) f.lf_quant[2 * plane + i] * (m.mul[2] * ep);
  out[1] = y;
  out[0] = x + m.kx * y;
  out[2] = b + m.kb * y;
}

// Final LF value of a cell: adaptive smoothing of interior cells (App. B.7) unless frame flag 0x80.
JXLB_HD void LfFinalCell(const FrameDev& f, const LfMul& m, uint32_t cx, uint32_t cy, float out[3]) {
  float c[3];
  LfDequantCell(f, m, cx, cy, c);
  const bool interior = cx > 0 && cy > 0 && cx + 1 < f.w8 && cy + 1 < f.h8;
  if ((f.flags & 0x80) || !interior) {
    out[0] = c[0];
    out[1] = c[1];
    out[2] = c[2];
    return;
  }
  const float w1 = 0.20345139757231578f, w2 = 0.0334829185968739f;
  const float w0 = 1.0f - 4.0f * (w1 + w2);
  float cross[3] = {0, 0, 0}, diag[3] = {0, 0, 0};
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      if (dx == 0 && dy == 0) continue;
      float v[3];
      LfDequantCell(f, m, cx + dx, cy + dy, v);
      float* acc = (dx == 0 || dy == 0) ? cross : diag;
      acc[0] += v[0];
      acc[1] += v[1];
      acc[2] += v[2];
    }
  float sm[3], gap = 0.5f;
  for (int k = 0; k < 3; ++k) {
    sm[k] = w0 * c[k] + w1 * cross[k] + w2 * diag[k];
    const float g = fabsf((c[k] - sm[k]) / m.mul[k]);
    if (g > gap) gap = g;
  }
  float fac = 3.0f - 4.0f * gap;
  if (fac < 0.0f) fac = 0.0f;
  for (int k = 0; k < 3; ++k) out[k] = (sm[k] - c[k]) * fac + c[k];
}

// ---- 64x64 regions -----------------------------------------------------------------------------------------------------
static constexpr int kRegionCells = 8;
static constexpr int kRegionDim = 64;
static constexpr int kTileStride = 65;                       // padded row stride (floats): conflict-free rows and columns
static constexpr int kTileFloats = kRegionDim * kTileStride;  // one channel

// Eight consecutive int16 coefficients (16-byte aligned in the plane), fetched with one 128-bit load on the device.
struct Coef8 {
  uint32_t w[4];
  JXLB_HD int Get(int j) const { return (int) (int16_t) (w[j >> 1] >> ((j & 1) * 16)); }
  JXLB_HD bool AllZero() const { return (w[0] | w[1] | w[2] | w[3]) == 0; }
};
JXLB_HD Coef8 LoadCoef8(const int16_t* p) {
  Coef8 c;
#ifdef __CUDA_ARCH__
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  c.w[0] = v.x;
  c.w[1] = v.y;
  c.w[2] = v.z;
  c.w[3] = v.w;
#else
  for (int i = 0; i < 4; ++i) c.w[i] = (uint32_t) (uint16_t) p[2 * i] | ((uint32_t) (uint16_t) p[2 * i + 1] << 16);
#endif
  return c;
}

struct RegionCell {
  uint8_t strategy;   // 0..26, 0xFF = outside the frame
  uint8_t flags;      // bit 0: this block is fully inside the region, bit 1: special 8x8 transform, bit 2: transposed layout
  int8_t ox, oy;      // block origin in region cell coordinates (may be negative when not contained)
  uint8_t kcols_log;  // log2 of the coefficient array's column count
  uint16_t hf_mul;
  uint32_t dq_off[3]; // offsets of the block's dequant matrices (X, Y, B) in the pool
  float scale;        // 65536 / global_scale / hf_mul
  float kx, kb;       // chroma-from-luma factors of the block
};

struct RegionShared {
  RegionCell cell[kRegionCells * kRegionCells];
  float lf[3][kRegionCells * kRegionCells];  // the region's LF samples (X, Y, B)
  float llf[1 + 4 + 16 + 64];                // LLF synthesis matrices A_1, A_2, A_4, A_8 (blocks inside a region span <= 8 cells)
  float tile[3 * kTileFloats];
};
JXLB_HD int LlfSharedOffset(int log2n) { return log2n == 0 ? 0 : log2n == 1 ? 1 : log2n == 2 ? 5 : 21; }

JXLB_HD float QmScale(uint32_t scale) {
  // 0.8 ^ (scale - 2)
  float v = 1.0f;
  if (scale >= 2)
    for (uint32_t i = 2; i < scale; ++i) v *= 0.8f;
  else
    for (uint32_t i = scale; i < 2; ++i) v *= 1.25f;
  return v;
}

// Dequantised coefficient F[v][u] (before CfL) of channel c at pixel offset (pr, pc) inside a block's rectangle.
JXLB_HD float DequantAt(const FrameDev& f, const NumericTables& nt, uint32_t t, uint32_t c, int q, uint32_t pr, uint32_t pc,
                        float block_scale) {
  if (q == 0) return 0.0f;
  const uint32_t cx = StrategyCellsX(t), cy = StrategyCellsY(t);
  const bool transposed = cy >= cx;
  const uint32_t kcols = 8 * (cx > cy ? cx : cy);
  const uint32_t kr = transposed ? pc : pr, kc = transposed ? pr : pc;
  const float w = nt.dequant[nt.dequant_off[StrategyQuantTable(t)][c] + kr * kcols + kc];
  return MulRn(AdjustQuantBiasRcp(q, c, nt.rcp11), MulRn(w, block_scale));
}

// Reconstructs the blocks fully contained in region (rx, ry) into f.xyb0.  sh: shared memory of the CTA.
template <class Sync>
JXLB_HD void ReconRegion(const FrameDev& f, const NumericTables& nt, uint32_t rx, uint32_t ry, RegionShared& sh, int tid,
                         int nthreads, Sync sync) {
  const uint32_t cx0 = rx * kRegionCells, cy0 = ry * kRegionCells;
  // P0: per-cell info
  for (int i = tid; i < kRegionCells * kRegionCells; i += nthreads) {
    const uint32_t ix = i % kRegionCells, iy = i / kRegionCells;
    RegionCell rc;
    rc.strategy = 0xFF;
    rc.flags = 0;
    rc.ox = rc.oy = 0;
    rc.hf_mul = 1;
    rc.kcols_log = 3;
    rc.dq_off[0] = rc.dq_off[1] = rc.dq_off[2] = 0;
    rc.scale = rc.kx = rc.kb = 0.0f;
    const uint32_t gx = cx0 + ix, gy = cy0 + iy;
    if (gx < f.w8 && gy < f.h8) {
      const size_t ci = (size_t) gy * f.w8 + gx;
      const uint32_t t = f.cell_strategy[ci] & 0x7F;
      const uint32_t off = f.cell_off[ci];
      const int dx = (int) (off & 0xFF), dy = (int) (off >> 8);
      rc.strategy = (uint8_t) t;
      rc.hf_mul = f.cell_hfmul[ci];
      rc.ox = (int8_t) ((int) ix - dx);
      rc.oy = (int8_t) ((int) iy - dy);
      const int bx = (int) StrategyCellsX(t), by = (int) StrategyCellsY(t);
      const bool contained = rc.ox >= 0 && rc.oy >= 0 && rc.ox + bx <= kRegionCells && rc.oy + by <= kRegionCells;
      const uint32_t qt = StrategyQuantTable(t);
      // bit 2: the dequant matrix must be read transposed (tall / square blocks whose matrix is not symmetric)
      rc.flags = (uint8_t) ((contained ? 1 : 0) | (IsSpecial8x8(t) ? 2 : 0) | ((by >= bx && !nt.dequant_symmetric[qt]) ? 4 : 0));
      rc.kcols_log = (uint8_t) (3 + FloorLog2((uint32_t) (bx > by ? bx : by)));
      for (int c = 0; c < 3; ++c) rc.dq_off[c] = nt.dequant_off[qt][c];
      rc.scale = 65536.0f / (float) f.global_scale / (float) rc.hf_mul;
      if (contained) {
        // CfL factors come from the 64x64 tile of the block's top-left corner
        const uint32_t tx = (cx0 + (uint32_t) rc.ox) / 8, ty = (cy0 + (uint32_t) rc.oy) / 8;
        rc.kx = f.cfl.base_x + (float) f.xfromy[(size_t) ty * f.w64 + tx] / (float) f.cfl.colour_factor;
        rc.kb = f.cfl.base_b + (float) f.bfromy[(size_t) ty * f.w64 + tx] / (float) f.cfl.colour_factor;
      }
    }
    sh.cell[i] = rc;
  }
  {
    const size_t lfplane0 = (size_t) f.h8 * f.lf_stride;
    for (int i = tid; i < 3 * kRegionCells * kRegionCells; i += nthreads) {
      const int c = i / (kRegionCells * kRegionCells), ci = i % (kRegionCells * kRegionCells);
      const uint32_t gx = cx0 + (uint32_t) (ci % kRegionCells), gy = cy0 + (uint32_t) (ci / kRegionCells);
      sh.lf[c][ci] = (gx < f.w8 && gy < f.h8) ? f.lf[c * lfplane0 + (size_t) gy * f.lf_stride + gx] : 0.0f;
    }
    for (int i = tid; i < 85; i += nthreads) {
      const int l = i < 1 ? 0 : i < 5 ? 1 : i < 21 ? 2 : 3;
      sh.llf[i] = nt.llf[l][i - LlfSharedOffset(l)];
    }
  }
  sync();
  // P1: dequantise + chroma-from-luma into the three tiles.  Work unit = 8 consecutive pixels of a row (one 8x8 cell, so
  // one RegionCell): three 128-bit coefficient loads per unit, all units of a thread issued before the first use so
  // that the DRAM latency is paid once per thread rather than once per pixel.
  const float xqm = QmScale(f.x_qm_scale), bqm = QmScale(f.b_qm_scale);
  const size_t cplane = (size_t) f.coef_h * f.coef_stride;
  constexpr int kUnits = kRegionDim * kRegionDim / 8, kRounds = 3;
  for (int base = tid; base < kUnits; base += kRounds * nthreads) {
    Coef8 raw[kRounds][3];
    RegionCell rcs[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      const int u = base + r * nthreads;
      if (u >= kUnits) break;
      const uint32_t row = (uint32_t) u / 8u, cg = (uint32_t) u % 8u;
      rcs[r] = sh.cell[(row / 8) * kRegionCells + cg];
      if (rcs[r].strategy != 0xFF && (rcs[r].flags & 1)) {
        const int16_t* p = f.coef + (size_t) (cy0 * 8 + row) * f.coef_stride + cx0 * 8 + cg * 8;
        raw[r][0] = LoadCoef8(p);
        raw[r][1] = LoadCoef8(p + cplane);
        raw[r][2] = LoadCoef8(p + 2 * cplane);
      }
    }
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
      const int u = base + r * nthreads;
      if (u >= kUnits) break;
      const uint32_t row = (uint32_t) u / 8u, col0 = ((uint32_t) u % 8u) * 8u;
      const RegionCell& rc = rcs[r];
      float* tx = sh.tile + (int) row * kTileStride + (int) col0;
      float* ty = tx + kTileFloats;
      float* tb = tx + 2 * kTileFloats;
      if (rc.strategy == 0xFF || !(rc.flags & 1)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) tx[j] = ty[j] = tb[j] = 0.0f;
        continue;
      }
      const uint32_t pr = row - (uint32_t) rc.oy * 8, pc0 = col0 - (uint32_t) rc.ox * 8;
      // index into the block's dequant matrix (stored like the coefficient array: transposed for square / tall blocks;
      // symmetric square matrices are read row-wise either way)
      const bool tr = (rc.flags & 4) != 0;
      const uint32_t k0 = tr ? ((pc0 << rc.kcols_log) + pr) : ((pr << rc.kcols_log) + pc0);
      const uint32_t kstep = tr ? (1u << rc.kcols_log) : 1u;
      const float sy = rc.scale, sx = rc.scale * xqm, sb = rc.scale * bqm;
      const bool zx = raw[r][0].AllZero(), zy = raw[r][1].AllZero(), zb = raw[r][2].AllZero();
      float vy[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int q = raw[r][1].Get(j);
        vy[j] = (!zy && q) ? MulRn(AdjustQuantBiasRcp(q, 1, nt.rcp11), MulRn(nt.dequant[rc.dq_off[1] + k0 + j * kstep], sy)) : 0.0f;
        ty[j] = vy[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int q = raw[r][0].Get(j);
        tx[j] = MulAddRn(rc.kx, vy[j], (!zx && q) ? MulRn(AdjustQuantBiasRcp(q, 0, nt.rcp11), MulRn(nt.dequant[rc.dq_off[0] + k0 + j * kstep], sx)) : 0.0f);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int q = raw[r][2].Get(j);
        tb[j] = MulAddRn(rc.kb, vy[j], (!zb && q) ? MulRn(AdjustQuantBiasRcp(q, 2, nt.rcp11), MulRn(nt.dequant[rc.dq_off[2] + k0 + j * kstep], sb)) : 0.0f);
      }
    }
  }
  sync();
  // P2: lowest frequencies from the (smoothed) LF image: one work item per (cell, channel); operands in shared memory
  for (int i = tid; i < 3 * kRegionCells * kRegionCells; i += nthreads) {
    const int c = i / (kRegionCells * kRegionCells), ci = i % (kRegionCells * kRegionCells);
    const int ix = ci % kRegionCells, iy = ci / kRegionCells;
    const RegionCell rc = sh.cell[ci];
    if (rc.strategy == 0xFF || !(rc.flags & 1)) continue;
    const int bx = (int) StrategyCellsX(rc.strategy), by = (int) StrategyCellsY(rc.strategy);
    const int kx = ix - rc.ox, ky = iy - rc.oy;
    const float* lf = sh.lf[c] + rc.oy * kRegionCells + rc.ox;
    float v;
    if (bx == 1 && by == 1) {
      v = lf[0];
    } else {
      const float* ay = sh.llf + LlfSharedOffset(FloorLog2((uint32_t) by)) + ky * by;
      const float* ax = sh.llf + LlfSharedOffset(FloorLog2((uint32_t) bx)) + kx * bx;
      v = 0.0f;
      for (int ny = 0; ny < by; ++ny) {
        float rowacc = 0.0f;
        for (int nx = 0; nx < bx; ++nx) rowacc += ax[nx] * lf[ny * kRegionCells + nx];
        v += ay[ny] * rowacc;
      }
    }
    sh.tile[c * kTileFloats + (rc.oy * 8 + ky) * kTileStride + rc.ox * 8 + kx] = v;
  }
  sync();
  // P3: special 8x8 transforms (disjoint from the DCT blocks) and the column pass of the DCT blocks
  for (int i = tid; i < 3 * kRegionCells * kRegionCells; i += nthreads) {
    const int c = i / (kRegionCells * kRegionCells), ci = i % (kRegionCells * kRegionCells);
    const RegionCell rc = sh.cell[ci];
    if (rc.strategy == 0xFF || (rc.flags & 3) != 3) continue;
    const int ix = ci % kRegionCells, iy = ci / kRegionCells;
    SpecialTransform8x8(rc.strategy, sh.tile + c * kTileFloats + iy * 8 * kTileStride + ix * 8, kTileStride, nt.afv_basis);
  }
  for (int i = tid; i < 3 * kRegionDim; i += nthreads) {
    const int c = i / kRegionDim, x = i % kRegionDim;
    for (int iy = 0; iy < kRegionCells;) {
      const RegionCell rc = sh.cell[iy * kRegionCells + x / 8];
      if (rc.strategy == 0xFF || (rc.flags & 3) != 1 || rc.oy != iy) {
        ++iy;
        continue;
      }
      const int by = (int) StrategyCellsY(rc.strategy);
      Idct1dDispatch(sh.tile + c * kTileFloats + iy * 8 * kTileStride + x, kTileStride, 8 * by);
      iy += by;
    }
  }
  sync();
  // P4: row pass
  for (int i = tid; i < 3 * kRegionDim; i += nthreads) {
    const int c = i / kRegionDim, y = i % kRegionDim;
    for (int ix = 0; ix < kRegionCells;) {
      const RegionCell rc = sh.cell[(y / 8) * kRegionCells + ix];
      if (rc.strategy == 0xFF || (rc.flags & 3) != 1 || rc.ox != ix) {
        ++ix;
        continue;
      }
      const int bx = (int) StrategyCellsX(rc.strategy);
      Idct1dDispatch(sh.tile + c * kTileFloats + y * kTileStride + ix * 8, 1, 8 * bx);
      ix += bx;
    }
  }
  sync();
  // P5: store
  const size_t pplane = (size_t) f.plane_h * f.plane_stride;
  for (int i = tid; i < kRegionDim * kRegionDim; i += nthreads) {
    const uint32_t col = i % kRegionDim, row = i / kRegionDim;
    const RegionCell rc = sh.cell[(row / 8) * kRegionCells + col / 8];
    if (rc.strategy == 0xFF || !(rc.flags & 1)) continue;
    const size_t gi = (size_t) (cy0 * 8 + row) * f.plane_stride + cx0 * 8 + col;
    const int ti = (int) row * kTileStride + (int) col;
    f.xyb0[gi] = sh.tile[ti];
    f.xyb0[pplane + gi] = sh.tile[kTileFloats + ti];
    f.xyb0[2 * pplane + gi] = sh.tile[2 * kTileFloats + ti];
  }
}

// Blocks that are not contained in one 64x64 region (larger than 64 pixels in a dimension, or straddling region
// borders): reconstructed in place in f.xyb0 through global memory.  (bx, by) = top-left cell of the block.
// Stage 1: dequantised coefficients + chroma-from-luma + the lowest frequencies from the LF image, written in the block's
// rectangle of f.xyb0 as F[vertical frequency][horizontal frequency].  The caller synchronises after it.
template <class Sync>
JXLB_HD void ReconLargeCoefficients(const FrameDev& f, const NumericTables& nt, uint32_t bx, uint32_t by, int tid, int nthreads,
                                    Sync sync) {
  const size_t ci = (size_t) by * f.w8 + bx;
  const uint32_t t = f.cell_strategy[ci] & 0x7F;
  const uint32_t cx = StrategyCellsX(t), cy = StrategyCellsY(t);
  const uint32_t R = 8 * cy, C = 8 * cx;
  const float inv_gs = 65536.0f / (float) f.global_scale;
  const float scale = inv_gs / (float) f.cell_hfmul[ci];
  const float xqm = QmScale(f.x_qm_scale), bqm = QmScale(f.b_qm_scale);
  const size_t cplane = (size_t) f.coef_h * f.coef_stride, pplane = (size_t) f.plane_h * f.plane_stride;
  const float kx = f.cfl.base_x + (float) f.xfromy[(size_t) (by / 8) * f.w64 + bx / 8] / (float) f.cfl.colour_factor;
  const float kb = f.cfl.base_b + (float) f.bfromy[(size_t) (by / 8) * f.w64 + bx / 8] / (float) f.cfl.colour_factor;
  for (uint32_t i = (uint32_t) tid; i < R * C; i += (uint32_t) nthreads) {
    const uint32_t pr = i / C, pc = i % C;
    const size_t gi = (size_t) (by * 8 + pr) * f.coef_stride + bx * 8 + pc;
    const float vy = DequantAt(f, nt, t, 1, f.coef[cplane + gi], pr, pc, scale);
    const float vx = DequantAt(f, nt, t, 0, f.coef[gi], pr, pc, scale * xqm) + kx * vy;
    const float vb = DequantAt(f, nt, t, 2, f.coef[2 * cplane + gi], pr, pc, scale * bqm) + kb * vy;
    const size_t po = (size_t) (by * 8 + pr) * f.plane_stride + bx * 8 + pc;
    f.xyb0[po] = vx;
    f.xyb0[pplane + po] = vy;
    f.xyb0[2 * pplane + po] = vb;
  }
  sync();
  const size_t lfplane = (size_t) f.h8 * f.lf_stride;
  for (uint32_t i = (uint32_t) tid; i < 3 * cx * cy; i += (uint32_t) nthreads) {
    const uint32_t c = i / (cx * cy), k = i % (cx * cy);
    const uint32_t kxi = k % cx, kyi = k / cx;
    const float* lf = f.lf + c * lfplane + (size_t) by * f.lf_stride + bx;
    const float* ay = nt.llf[FloorLog2(cy)] + kyi * cy;
    const float* ax = nt.llf[FloorLog2(cx)] + kxi * cx;
    float v = 0.0f;
    for (uint32_t ny = 0; ny < cy; ++ny) {
      float rowacc = 0.0f;
      for (uint32_t nx = 0; nx < cx; ++nx) rowacc += ax[nx] * lf[(size_t) ny * f.lf_stride + nx];
      v += ay[ny] * rowacc;
    }
    f.xyb0[c * pplane + (size_t) (by * 8 + kyi) * f.plane_stride + bx * 8 + kxi] = v;
  }
}

// Stage 2: separable inverse DCT in place (columns, then rows).
template <class Sync>
JXLB_HD void ReconLargeInverse(const FrameDev& f, uint32_t bx, uint32_t by, int tid, int nthreads, Sync sync) {
  const uint32_t t = f.cell_strategy[(size_t) by * f.w8 + bx] & 0x7F;
  const uint32_t R = 8 * StrategyCellsY(t), C = 8 * StrategyCellsX(t);
  const size_t pplane = (size_t) f.plane_h * f.plane_stride;
  float scratch[512];
  for (uint32_t i = (uint32_t) tid; i < 3 * C; i += (uint32_t) nthreads) {
    const uint32_t c = i / C, u = i % C;
    float* p = f.xyb0 + c * pplane + (size_t) (by * 8) * f.plane_stride + bx * 8 + u;
    if (R > 64) Idct1dLarge(p, (int) f.plane_stride, (int) R, scratch);
    else Idct1dDispatch(p, (int) f.plane_stride, (int) R);
  }
  sync();
  for (uint32_t i = (uint32_t) tid; i < 3 * R; i += (uint32_t) nthreads) {
    const uint32_t c = i / R, v = i % R;
    float* p = f.xyb0 + c * pplane + (size_t) (by * 8 + v) * f.plane_stride + bx * 8;
    if (C > 64) Idct1dLarge(p, 1, (int) C, scratch);
    else Idct1dDispatch(p, 1, (int) C);
  }
}

// Blocks that are not contained in one 64x64 region (larger than 64 pixels in a dimension, or straddling region
// borders): reconstructed in place in f.xyb0 through global memory.  (bx, by) = top-left cell of the block.
template <class Sync>
JXLB_HD void ReconLargeBlock(const FrameDev& f, const NumericTables& nt, uint32_t bx, uint32_t by, int tid, int nthreads,
                             Sync sync) {
  ReconLargeCoefficients(f, nt, bx, by, tid, nthreads, sync);
  sync();
  ReconLargeInverse(f, bx, by, tid, nthreads, sync);
}

// True when the block whose top-left cell is (bx, by) is handled by ReconLargeBlock rather than ReconRegion.
JXLB_HD bool BlockNeedsLargePath(uint32_t t, uint32_t bx, uint32_t by) {
  const uint32_t cx = StrategyCellsX(t), cy = StrategyCellsY(t);
  return (bx % kRegionCells) + cx > (uint32_t) kRegionCells || (by % kRegionCells) + cy > (uint32_t) kRegionCells;
}

}  // namespace jxlb
