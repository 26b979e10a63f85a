"""GPU parity on files whose modular streams carry palette transforms (tests/palette_cases.py): the CUDA path through the
C ABI against the reference run on the same file -- bit-exact for every lossless file in every output configuration the
reference offers, exact alpha + the lossy bound on colour for the VarDCT files with a lossless alpha."""
import numpy as np
import pytest

import golden_lib
import palette_cases as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def J():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import jxl_coder_b200 as J
    J.load_library()
    return J


@pytest.mark.parametrize("name", sorted(P.CASES))
def test_palette_files_match_reference(J, ref, name):
    data, img, bits, kw = P.make(ref, name)
    for cfg in ((2, 3) if bits == 8 else (2, 3, 5)):
        want = ref.decode_sampled(data, cfg=cfg)
        got = J.JxlCoder.decode(data, cfg)
        assert (got.width, got.height) == (want["width"], want["height"])
        a, b = got.pixels, want["pixels"]
        assert a.shape == b.shape, (name, cfg)
        if kw.get("lossless"):
            assert np.array_equal(a, b), (name, cfg, int((a != b).sum()))
        elif cfg == 2:
            h, w = img.shape[:2]
            a4, b4 = a[:, : w * 4].reshape(h, w, 4), b[:, : w * 4].reshape(h, w, 4)
            assert np.array_equal(a4[..., 3], b4[..., 3]), name
            golden_lib.lossy_close(a4, b4, name)


def test_palette_batch_mixes_with_other_images(J, ref):
    """Palette and non-palette images in one batch (different channel plans per image in the same kernels)."""
    import cases
    names = ["pal_rgb_12colours_700x300", "pal_local_rgba_822x769", "pal_16bit_x257_455x482"]
    datas = [P.make(ref, n)[0] for n in names] + [cases.get(cases.SMALL[0])]
    out = J.decode_batch(datas, config=2)
    for d, o in zip(datas, out):
        want = ref.decode_sampled(d, cfg=2)["pixels"]
        if d is datas[-1] and "lossless" not in cases.SMALL[0]:
            golden_lib.lossy_close(o.pixels, want, "batch")
        else:
            assert np.array_equal(o.pixels, want)
