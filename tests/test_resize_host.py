"""Pins the rescale restatement (csrc/resize.cc: dimension logic of weaver/src/scale.rs + pic-scale 0.7.6's fixed-point
resampler) against the reference binary: lossless inputs, so the decode is exact and RescaleImage is what is compared.
Bit-exact on every supported case."""
import numpy as np
import pytest

import cases
import hostemu_lib as H


def _image(w, h, seed):
    rng = np.random.default_rng(seed)
    base = np.cumsum(rng.standard_normal((h, w, 3)), axis=1) * 6 + 128 + rng.standard_normal((h, w, 3)) * 20
    img = np.empty((h, w, 4), np.uint8)
    img[..., :3] = np.clip(base, 0, 255).astype(np.uint8)
    img[..., 3] = 255
    return img


CASES = [  # (w, h, req_w, req_h, scale_mode, filter)
    (96, 64, 24, 16, 3, 4), (96, 64, 35, 23, 3, 4), (96, 64, 12, 8, 3, 1), (96, 64, 96, 20, 3, 6), (96, 64, 31, 64, 3, 7),
    (120, 90, 40, 40, 1, 4), (120, 90, 40, 30, 2, 4), (120, 90, 60, -1, 1, 6), (120, 90, -2, 33, 3, 1), (120, 90, 50, -2, 1, 7),
    (200, 120, 50, 30, 1, 4), (200, 120, 67, 41, 3, 6), (64, 200, 20, 100, 1, 1),
]


@pytest.mark.parametrize("case", CASES)
def test_resize_matches_reference(case, ref):
    w, h, rw, rh, mode, filt = case
    img = _image(w, h, w * 1000 + h)
    data = cases._cached("resize_src_%dx%d" % (w, h), lambda: ref.encode(img[..., :3].reshape(-1), w, h, colorspace=1, compression=1))
    r = ref.decode_sampled(data, w=rw, h=rh, cfg=2, scale_mode=mode, filt=filt)
    want = r["pixels"][:, : r["width"] * 4].reshape(r["height"], r["width"], 4)
    got = H.resize_rgba8(img, rw, rh, mode, filt)
    assert not isinstance(got, int), "plan status %s" % got
    assert got.shape == want.shape, (got.shape, want.shape)
    assert (got == want).all(), (np.abs(got.astype(int) - want.astype(int)).max(), (got != want).mean())


@pytest.mark.parametrize("args", [(96, 64, 200, 64, 3, 4), (96, 64, 24, 16, 3, 5), (96, 64, 24, 16, 3, 2), (96, 64, 24, 16, 3, 8), (120, 90, 40, 40, 2, 4),
                                  (120, 90, 100, 20, 2, 1), (120, 90, 50, -2, 2, 7)])
def test_unpinned_cases_are_refused(args):
    """Upscaling, ScaleToFill crops and the filters whose pic-scale arithmetic is not pinned yet must be refused, not
    approximated."""
    w, h, rw, rh, mode, filt = args
    assert H.resize_rgba8(_image(w, h, 1), rw, rh, mode, filt) == 1
