"""GPU parity on frames coded at half resolution (tests/upsampling_cases.py) through the C ABI: one batch, then one of
them through decodeSampled (rescale + 1010102) and with an orientation."""
import numpy as np
import pytest

import golden_lib
import upsampling_cases as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def J():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import jxl_coder_b200 as J
    J.load_library()
    return J


def test_upsampled_files_match_reference(J, ref):
    datas = [U.make(ref, *g) for g in U.GRID]
    outs = J.decode_batch(datas, config=2)
    for g, d, o in zip(U.GRID, datas, outs):
        want = U.ref_decode_stable(ref, d, cfg=2)["pixels"]
        assert o.pixels.shape == want.shape, g
        if g[5] is None:
            golden_lib.lossy_close(o.pixels, want, U.name(*g), min_exact=0.97)
        else:
            # an upsampled alpha is float work in the reference too (within one step), and the output is PREMULTIPLIED: one
            # step in alpha times one step in colour reaches 2 on isolated samples (seen: 1 sample in 1.5 million)
            d = np.abs(o.pixels.astype(np.int32) - want.astype(np.int32))
            assert d.max() <= 2 and float((d > 1).mean()) < 1e-5 and float((d == 0).mean()) >= 0.97, (g, int(d.max()), float((d > 1).mean()))


def test_upsampled_file_through_decode_sampled(J, ref):
    d = U.make(ref, 600, 400, 12.0, -1, 7, None)
    want = ref.decode_sampled(d, w=300, h=200, cfg=2, scale_mode=1, filt=1)
    got = J.JxlCoder.decode_sampled(d, 300, 200, 2, 1, 1)
    assert (got.width, got.height) == (want["width"], want["height"])
    dd = np.abs(got.pixels.astype(int) - want["pixels"].astype(int))
    assert dd.max() <= 1
    want = ref.decode_sampled(d, cfg=3)  # RGBA_F16
    got = J.JxlCoder.decode(d, 3)
    fa = got.pixels.view(np.float16).astype(np.float32)
    fb = want["pixels"].view(np.float16).astype(np.float32)
    assert np.abs(fa - fb).max() <= 1.25 / 255  # one 8-bit step plus the half-float rounding near 1.0 (2^-11)
