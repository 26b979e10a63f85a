"""TEST/BUILD INFRASTRUCTURE — regenerates the two constant tables the product embeds, from the reference's shipped
libjxl 0.12.0 binary (run where oracle/_ref was built):
  * jxl_coder_b200/csrc/tables/dither_table.inc : the 32x32 blue-noise table added before 8-bit rounding
    (file offset 0x1c300 of lib/x86_64/libjxl.so),
  * jxl_coder_b200/csrc/tables/afv_basis.inc    : the 16x16 AFV basis of the JPEG XL format (rows 1..15 read by
    pushing unit coefficients through the binary's own TransformToPixels; row 0 is the constant 0.25).
These are data constants of the codec/format, not code."""
import ctypes as C
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import refjxl  # noqa: E402

OUT = os.path.join(HERE, '..', 'jxl_coder_b200', 'csrc', 'tables')


def _lit(v):
    s = '%.9g' % v
    if '.' not in s and 'e' not in s and 'n' not in s:
        s += '.0'
    return s + 'f'


def fmt(vals, per_line=8):
    lines = []
    for i in range(0, len(vals), per_line):
        lines.append(', '.join(_lit(v) for v in vals[i:i + per_line]) + ',')
    return '\n'.join(lines) + '\n'


def main():
    os.makedirs(OUT, exist_ok=True)
    d = refjxl.dither_table().astype(np.float32).ravel()
    open(os.path.join(OUT, 'dither_table.inc'), 'w').write(fmt(d))
    L = refjxl.lib()
    L.ref_transform_to_pixels.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    B = np.zeros((16, 16), np.float32)
    for iy in range(4):
        for ix in range(4):
            K = np.zeros((8, 8), np.float32)
            K[2 * iy, 2 * ix] = 1
            px = np.zeros((8, 8), np.float32)
            assert L.ref_transform_to_pixels(14, K.ctypes.data, 64, px.ctypes.data, 8) == 0
            B[iy * 4 + ix] = px[:4, :4].ravel()
    B[0] = 0.25
    assert np.abs(B.astype(np.float64) @ B.astype(np.float64).T - np.eye(16)).max() < 1e-6
    open(os.path.join(OUT, 'afv_basis.inc'), 'w').write(fmt(B.ravel()))
    print('wrote', OUT)


if __name__ == '__main__':
    main()
