"""Pins the orientation mapping used by OrientKernel (csrc/kernels_post.cu) against the reference: libjxl applies
the codestream orientation to the pixels DecodeJpegXlOneShot receives (interop/JxlDecoding.cpp:36-176 never sets
keep_orientation).  `orient` below is the same table in numpy."""
import numpy as np
import pytest

import cases


def orient(img, o):
    """img [h,w,...] in coded orientation -> displayed orientation o (EXIF numbering)."""
    if o == 2:
        return img[:, ::-1]
    if o == 3:
        return img[::-1, ::-1]
    if o == 4:
        return img[::-1]
    if o == 5:
        return img.swapaxes(0, 1)
    if o == 6:
        return img[::-1].swapaxes(0, 1)
    if o == 7:
        return img[::-1, ::-1].swapaxes(0, 1)
    if o == 8:
        return img[:, ::-1].swapaxes(0, 1)
    return img


def source(alpha=False, w=72, h=40):
    rng = np.random.default_rng(9)
    return rng.integers(0, 256, (h, w, 4 if alpha else 3)).astype(np.uint8)


def encoded(ref, o, alpha=False):
    img = source(alpha)
    h, w, c = img.shape
    return img, cases._cached("orient_%d_%d" % (o, int(alpha)), lambda: ref.encode_ex(img.reshape(-1), w, h, c, lossless=True, orientation=o))


@pytest.mark.parametrize("o", range(1, 9))
def test_orientation_table_matches_reference(o, ref):
    img, data = encoded(ref, o)
    r = ref.decode_sampled(data, cfg=2)
    want = r["pixels"][:, : r["width"] * 4].reshape(r["height"], r["width"], 4)
    mine = orient(img, o)
    assert want.shape[:2] == mine.shape[:2]
    assert (want[..., :3] == mine).all()
    assert ref.get_size(data) == (mine.shape[1], mine.shape[0])


@pytest.mark.parametrize("o", range(2, 9))
def test_lossy_dither_follows_orientation(o, ref):
    """Lossy (XYB, 8-bit dithered) frames: the CPU emulation of the kernel code, oriented, against the reference.  Pins
    DitherIndex (csrc/pixel_stages.h): flips happen before libjxl's dither, the transposition after it."""
    import hostemu_lib as H
    from oracle import synth
    import golden_lib
    w, h = 200, 136
    img = synth.synth_image(w, h, 5)
    data = cases._cached("orient%d_lossy" % o, lambda: ref.encode_ex(img, w, h, 3, distance=1.0, orientation=o))
    r = ref.decode_sampled(data, cfg=2)
    want = r["pixels"][:, : r["width"] * 4].reshape(r["height"], r["width"], 4)
    e = H.Decoded(data)
    emu = e.render()
    e.close()
    golden_lib.lossy_close(orient(emu, o), want, "orientation %d" % o)
