"""Host-side hardening of the header parser: Lehmer-coded permutations in O(n log n), frame sizes from
the bitstream bounded before anything is allocated, features that are parsed but not decoded are refused."""
import time

import numpy as np
import pytest

import hostemu_lib as H


def naive_lehmer(size, digits):
    rest = list(range(size))
    out = []
    for i in range(size):
        k = digits[i] if i < len(digits) else 0
        out.append(rest.pop(k))
    return out


@pytest.mark.parametrize("size,end,seed", [(1, 0, 0), (1, 1, 0), (2, 2, 1), (64, 64, 2), (64, 10, 3), (1000, 1000, 4), (4096, 4096, 5), (777, 300, 6)])
def test_lehmer_matches_naive(size, end, seed):
    rng = np.random.default_rng(seed)
    digits = np.zeros(size, np.uint32)
    for i in range(end):
        digits[i] = rng.integers(0, size - i)
    want = naive_lehmer(size, list(digits[:end]))
    perm = np.zeros(size, np.uint32)
    d = digits.copy()
    assert H.lib().emu_expand_lehmer(size, end, perm.ctypes.data, d.ctypes.data) == 0
    assert perm.tolist() == want


def test_lehmer_worst_case_is_fast():
    # every digit maximal: the shifting-list decode moves ~size^2 / 2 = 2e9 entries here; the tree takes milliseconds
    size = 65536
    digits = np.arange(size - 1, -1, -1, dtype=np.uint32)
    perm = np.zeros(size, np.uint32)
    t0 = time.time()
    assert H.lib().emu_expand_lehmer(size, size, perm.ctypes.data, digits.ctypes.data) == 0
    assert time.time() - t0 < 0.5
    assert perm.tolist() == list(range(size - 1, -1, -1))


class Bits:
    """JPEG XL bit order: first bit read is the LSB of the first byte."""

    def __init__(self):
        self.v, self.n = 0, 0

    def put(self, value, nbits):
        self.v |= (value & ((1 << nbits) - 1)) << self.n
        self.n += nbits

    def u32(self, selector, value=0, nbits=0):
        self.put(selector, 2)
        self.put(value, nbits)

    def bytes(self):
        return self.v.to_bytes((self.n + 7) // 8, "little")


def tiny_file_with_crop(w_extra_bits_value, h_extra_bits_value):
    b = Bits()
    b.put(0x0AFF, 16)          # signature FF 0A
    b.put(1, 1)                # SizeHeader: small
    b.put(0, 5)                # ysize = 8
    b.put(1, 3)                # ratio 1:1 -> xsize = 8
    b.put(1, 1)                # ImageMetadata all_default
    b.put(1, 1)                # default_transform
    b.put(0, 5)                # pad to the byte boundary (frame headers are byte aligned)
    b.put(0, 1)                # FrameHeader all_default = 0
    b.put(0, 2)                # frame_type regular
    b.put(1, 1)                # encoding modular
    b.u32(0)                   # flags = 0
    b.u32(0)                   # upsampling 1
    b.put(1, 2)                # group_size_shift
    b.u32(0)                   # num_passes 1
    b.put(1, 1)                # have_crop
    b.u32(0, 0, 8)             # x0 = 0
    b.u32(0, 0, 8)             # y0 = 0
    b.u32(3, w_extra_bits_value, 30)   # width  = 18688 + value
    b.u32(3, h_extra_bits_value, 30)   # height = 18688 + value
    b.u32(0)                   # blend mode replace (crop covers the canvas: no source field)
    b.put(1, 1)                # is_last
    b.u32(0)                   # name length 0
    b.put(1, 1)                # restoration filter all_default
    b.u32(0)                   # extensions
    b.put(0, 7)
    return b.bytes() + bytes(40)


def test_oversized_crop_frame_is_refused_at_once():
    import jxl_coder_b200 as J
    data = tiny_file_with_crop((1 << 30) - 1, 65280 - 18688)
    t0 = time.time()
    with pytest.raises(J.JxlCoderError) as e:
        J.JxlCoder.decode(data)
    assert time.time() - t0 < 1.0
    assert e.value.status == 1, (e.value.status, e.value.message)   # JXLB_INVALID_JXL
    assert "frame size" in e.value.message or "TOC" in e.value.message
    # the animated-image constructor walks the same frame headers
    with pytest.raises(J.JxlCoderError):
        J.JxlAnimatedImage(data)
