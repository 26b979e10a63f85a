"""Progressive passes through the CPU emulation of the kernel code, against the reference's decode."""
import pytest

import golden_lib
import hostemu_lib as H
import progressive_cases as P


@pytest.mark.parametrize("w,h,opt,dist,effort", P.GRID)
def test_progressive_passes(w, h, opt, dist, effort, ref):
    data = P.make(ref, w, h, opt, dist, effort)
    want = ref.decode_sampled(data, cfg=2)["pixels"][:, : w * 4].reshape(h, w, 4)
    e = H.Decoded(data)
    assert e.status == 0
    out = e.render()
    e.close()
    golden_lib.lossy_close(out, want, P.name(w, h, opt, dist, effort))
