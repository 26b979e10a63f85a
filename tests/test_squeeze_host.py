"""Squeezed (lossy) alpha channels -- csrc/squeeze.h, the global-stream decode in frame_parser.cc and the per-group
channel rectangles in vardct_sections.h -- through the CPU emulation of the kernel code, against the reference:
the alpha plane is integer work and must be bit-exact; colour stays within the lossy bound."""
import numpy as np
import pytest

import cases
import golden_lib
import hostemu_lib as H


def rgba_lossy(ref, w, h, seed, alpha_distance=1.0):
    from oracle import synth
    img = synth.synth_image(w, h, seed, alpha=True)
    return cases._cached("rgba_lossy_sq_%dx%d_s%d_a%g" % (w, h, seed, alpha_distance),
                         lambda: ref.encode_ex(img, w, h, 4, distance=1.0, alpha_distance=alpha_distance))


# the first is a single-section frame (the whole pyramid is in the global stream); the last two are wider / taller than 2048: part of the pyramid is coded in the LF groups (ModularLfGroup)
@pytest.mark.parametrize("shape", [(120, 90, 40, 1.0), (320, 264, 41, 1.0), (257, 300, 42, 2.0), (700, 520, 43, 0.5), (1024, 300, 44, 1.0), (2304, 300, 45, 1.0), (520, 4200, 46, 1.0)])
def test_squeezed_alpha_matches_reference(shape, ref):
    w, h, seed, ad = shape
    data = rgba_lossy(ref, w, h, seed, ad)
    r = ref.decode_sampled(data, cfg=2)
    want = r["pixels"][:, : w * 4].reshape(h, w, 4)
    e = H.Decoded(data)
    emu = e.render()
    e.close()
    assert (emu[..., 3] == want[..., 3]).all()
    assert want[..., 3].min() < 255
    a = emu[..., 3:4].astype(np.uint16)
    emu[..., :3] = (emu[..., :3].astype(np.uint16) * a // 255).astype(np.uint8)  # ReformatColorConfig premultiplies
    golden_lib.lossy_close(emu, want, "squeezed alpha %dx%d" % (w, h))


def test_squeeze_layout_default_parameters():
    """1 channel of 1000 x 600 (wide): horizontal first; the channel list is sorted small to large and partitions the plane."""
    lib = H.lib()
    assert lib is not None
    # checked through the decoder above; here only the arithmetic identity of the pyramid sizes
    w, h = 1000, 600
    sizes = []
    cw, ch = w, h
    wide = w > h
    if not wide and ch > 8:
        sizes.append((cw, ch - (ch + 1) // 2))
        ch = (ch + 1) // 2
    while cw > 8 or ch > 8:
        if cw > 8:
            sizes.append((cw - (cw + 1) // 2, ch))
            cw = (cw + 1) // 2
        if ch > 8:
            sizes.append((cw, ch - (ch + 1) // 2))
            ch = (ch + 1) // 2
    assert sum(a * b for a, b in sizes) + cw * ch == w * h
