"""Progressive passes through the CPU emulation of the kernel code, against the reference's decode."""
import numpy as np
import pytest

import golden_lib
import hostemu_lib as H
import progressive_cases as P


@pytest.mark.parametrize("w,h,opt,dist,effort,ad", P.GRID)
def test_progressive_passes(w, h, opt, dist, effort, ad, ref):
    data = P.make(ref, w, h, opt, dist, effort, ad)
    want = ref.decode_sampled(data, cfg=2)["pixels"][:, : w * 4].reshape(h, w, 4)
    e = H.Decoded(data)
    assert e.status == 0
    out = e.render()
    e.close()
    if ad is not None:
        assert np.array_equal(out[..., 3], want[..., 3])  # alpha is integer work: exact
        a = out[..., 3:4].astype(np.uint16)
        out[..., :3] = (out[..., :3].astype(np.uint16) * a // 255).astype(np.uint8)  # ReformatColorConfig premultiplies
    golden_lib.lossy_close(out, want, P.name(w, h, opt, dist, effort, ad))
