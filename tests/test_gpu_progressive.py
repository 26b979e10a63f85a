"""GPU parity on progressive VarDCT files (tests/progressive_cases.py) through the C ABI, one batch."""
import pytest

import golden_lib
import progressive_cases as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def J():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import jxl_coder_b200 as J
    J.load_library()
    return J


def test_progressive_files_match_reference(J, ref):
    datas = [P.make(ref, *g) for g in P.GRID]
    outs = J.decode_batch(datas, config=2)
    for g, d, o in zip(P.GRID, datas, outs):
        want = ref.decode_sampled(d, cfg=2)["pixels"]
        assert o.pixels.shape == want.shape, g
        golden_lib.lossy_close(o.pixels, want, P.name(*g))
