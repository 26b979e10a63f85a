"""GPU parity over the option space of the reference's own encoder (tests/encoder_space.py) through the C ABI."""
import numpy as np
import pytest

import encoder_space as E
import golden_lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def J():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import jxl_coder_b200 as J
    J.load_library()
    return J


def test_reference_encoder_space_batch(J, ref):
    datas = [E.make(ref, *g) for g in E.GRID]
    outs = J.decode_batch(datas, config=2)
    for g, d, o in zip(E.GRID, datas, outs):
        want = ref.decode_sampled(d, cfg=2)["pixels"]
        assert o.pixels.shape == want.shape, g
        if g[0]:
            assert np.array_equal(o.pixels, want), g
        else:
            golden_lib.lossy_close(o.pixels, want, E.name(*g), min_exact=0.97)
