"""Files made with every (lossless | lossy, effort, decoding speed) the reference's encoder offers, through the CPU emulation
of the kernel code against the reference's decode."""
import numpy as np
import pytest

import encoder_space as E
import golden_lib
import hostemu_lib as H


@pytest.mark.parametrize("lossless,effort,ds", E.GRID)
def test_reference_encoder_space(lossless, effort, ds, ref):
    data = E.make(ref, lossless, effort, ds)
    want = ref.decode_sampled(data, cfg=2)["pixels"][:, : E.W * 4].reshape(E.H, E.W, 4)
    e = H.Decoded(data)
    assert e.status == 0
    out = e.render()
    e.close()
    a = out[..., 3:4].astype(np.uint16)
    out[..., :3] = (out[..., :3].astype(np.uint16) * a // 255).astype(np.uint8)  # ReformatColorConfig premultiplies
    if lossless:
        assert np.array_equal(out, want)
    else:
        golden_lib.lossy_close(out, want, E.name(lossless, effort, ds), min_exact=0.97)
