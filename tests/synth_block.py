"""Synthetic single-block inputs for the reconstruction code (csrc/test_block.h): sparse quantised coefficients with a
realistic decay towards high frequencies, plus LF samples.  Plain DCT strategies only (8x8 ... 256x256)."""
import numpy as np

CELLS_X = [1, 1, 1, 1, 2, 4, 1, 2, 1, 4, 2, 4, 1, 1, 1, 1, 1, 1, 8, 4, 8, 16, 8, 16, 32, 16, 32]
CELLS_Y = [1, 1, 1, 1, 2, 4, 2, 1, 4, 1, 4, 2, 1, 1, 1, 1, 1, 1, 8, 8, 4, 16, 16, 8, 32, 32, 16]
PLAIN_DCT = [0, 4, 5, 6, 7, 8, 9, 10, 11, 18, 19, 20, 21, 22, 23, 24, 25, 26]
NAMES = {0: "DCT8", 4: "DCT16", 5: "DCT32", 6: "DCT16x8", 7: "DCT8x16", 8: "DCT32x8", 9: "DCT8x32", 10: "DCT32x16", 11: "DCT16x32",
         18: "DCT64", 19: "DCT64x32", 20: "DCT32x64", 21: "DCT128", 22: "DCT128x64", 23: "DCT64x128", 24: "DCT256", 25: "DCT256x128",
         26: "DCT128x256"}


def make(strategy, seed, dense=False):
    cx, cy = CELLS_X[strategy], CELLS_Y[strategy]
    R, C = 8 * cy, 8 * cx
    rng = np.random.default_rng(seed * 31 + strategy)
    v, u = np.mgrid[0:R, 0:C]
    decay = 1.0 / (1.0 + 0.15 * (v * 64.0 / R + u * 64.0 / C))
    q = np.rint(rng.standard_normal((3, R, C)) * 40.0 * decay).astype(np.int16)
    if not dense:
        q[rng.random((3, R, C)) < 0.6] = 0
    q[:, :cy, :cx] = 0  # the lowest frequencies come from the LF image
    lf = (rng.standard_normal((3, cy, cx)) * 0.2).astype(np.float32)
    return np.ascontiguousarray(q), np.ascontiguousarray(lf)


def libjxl_layout(strategy, plane):
    """[vertical][horizontal] coefficient plane -> libjxl's block array (8 min x 8 max, transposed for tall / square blocks)."""
    return np.ascontiguousarray(plane.T if CELLS_Y[strategy] >= CELLS_X[strategy] else plane)
