"""2x upsampled frames through the CPU emulation of the kernel code, against the reference's decode."""
import pytest

import golden_lib
import hostemu_lib as H
import upsampling_cases as U


@pytest.mark.parametrize("w,h,dist,res,effort,ad", U.GRID)
def test_upsampled_frames(w, h, dist, res, effort, ad, ref):
    data = U.make(ref, w, h, dist, res, effort, ad)
    want = U.ref_decode_stable(ref, data, cfg=2)["pixels"][:, : w * 4].reshape(h, w, 4)
    e = H.Decoded(data)
    assert e.status == 0 and e.info["upsampling"] == 2
    out = e.render()
    e.close()
    if ad is not None:  # the upsampled alpha is float work in the reference too: within one step; then premultiply with the reference's alpha
        import numpy as np
        assert np.abs(out[..., 3].astype(int) - want[..., 3].astype(int)).max() <= 1
        a = want[..., 3:4].astype(np.uint16)
        out[..., :3] = (out[..., :3].astype(np.uint16) * a // 255).astype(np.uint8)
    golden_lib.lossy_close(out, want, U.name(w, h, dist, res, effort, ad), min_exact=0.97)
