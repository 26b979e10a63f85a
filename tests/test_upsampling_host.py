"""2x upsampled frames through the CPU emulation of the kernel code, against the reference's decode."""
import pytest

import golden_lib
import hostemu_lib as H
import upsampling_cases as U


@pytest.mark.parametrize("w,h,dist,res,effort", U.GRID)
def test_upsampled_frames(w, h, dist, res, effort, ref):
    data = U.make(ref, w, h, dist, res, effort)
    want = ref.decode_sampled(data, cfg=2)["pixels"][:, : w * 4].reshape(h, w, 4)
    e = H.Decoded(data)
    assert e.status == 0 and e.info["upsampling"] == 2
    out = e.render()
    e.close()
    golden_lib.lossy_close(out, want, U.name(w, h, dist, res, effort))
