"""TEST INFRASTRUCTURE — deterministic synthetic images of SURVEY.md §8d and their encoding with the reference's own
encoder settings (EncodeJxlOneshot, /root/reference/jxlcoder/src/main/cpp/interop/JxlEncoding.cpp:48-193:
quality 90 ⇒ distance 1.0, effort 7, decoding speed 0, sRGB).  Used by tests/ and by bench.py to build inputs."""
import numpy as np


def synth_image(w, h, index=0, alpha=False):
    """smooth sinusoid field + 64-px checker ±20 + LCG noise ±8 (LCG s = s*1664525 + 1013904223, seed 12345+index);
    alpha = horizontal ramp 255→128."""
    n = w * h
    a = np.uint32(1664525)
    c = np.uint32(1013904223)
    # s_k = a^k s_0 + c (a^(k-1) + ... + 1)  (mod 2^32), k = 1..n
    with np.errstate(over='ignore'):
        ak = np.cumprod(np.full(n, a, dtype=np.uint32), dtype=np.uint32)            # a^1..a^n
        geo = np.concatenate([[np.uint32(1)], ak[:-1]]).astype(np.uint32)          # a^0..a^(n-1)
        gsum = np.cumsum(geo, dtype=np.uint32)                                      # sum_{j<k} a^j
        s = ak * np.uint32((12345 + index) & 0xFFFFFFFF) + c * gsum
    nz = ((s >> np.uint32(24)) & np.uint32(15)).astype(np.int32).reshape(h, w) - 8
    y, x = np.mgrid[0:h, 0:w]
    r = 127.5 + 100 * np.sin(x * 0.013) * np.cos(y * 0.007)
    g = 127.5 + 100 * np.sin((x + y) * 0.005)
    b = 127.5 + 100 * np.cos(x * 0.002 - y * 0.011)
    blk = np.where(((x >> 6) ^ (y >> 6)) & 1, 20, -20)
    ch = [r.astype(np.int32) + nz + blk, g.astype(np.int32) + nz, b.astype(np.int32) - nz]
    if alpha:
        ch.append(255 - ((x * 255) // w) // 2)
    return np.clip(np.stack(ch, -1), 0, 255).astype(np.uint8)
