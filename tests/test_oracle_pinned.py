"""Pins the oracle: (1) the committed golden vectors are what the reference (oracle/_ref) produces today, (2) the
Python restatement (oracle/pyjxl) reproduces the reference on them, (3) every inverse transform of the restatement
matches the reference binary's own TransformToPixels."""
import ctypes as C

import numpy as np
import pytest

import golden_lib
from oracle.pyjxl import decode as D
from oracle.pyjxl import vardct as vd


@pytest.mark.parametrize("name", golden_lib.names())
def test_golden_matches_reference(name, ref):
    g = golden_lib.load(name)
    raw, _ = ref.decode_oneshot(g["jxl"])
    assert (raw == g["raw"]).all()
    for k in g:
        if k.startswith("out_"):
            cfg = int(k[4:])
            r = ref.decode_sampled(g["jxl"], cfg=cfg)
            assert (r["pixels"] == g[k]).all()


@pytest.mark.parametrize("name", golden_lib.names())
def test_restatement_matches_golden(name):
    g = golden_lib.load(name)
    r = D.decode(g["jxl"])
    if "lossless" in name:
        assert (r["rgba"] == g["raw"]).all()
    else:
        golden_lib.restatement_close(r["rgba"], g["raw"], name)


def test_inverse_transforms_match_reference_binary(ref):
    L = ref.lib()
    L.ref_transform_to_pixels.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    rng = np.random.default_rng(0)
    for t in range(27):
        cx, cy = vd.CBX[t], vd.CBY[t]
        kr, kc = 8 * min(cx, cy), 8 * max(cx, cy)
        K = rng.standard_normal((kr, kc))
        Kf = np.ascontiguousarray(K, np.float32)
        px = np.zeros((8 * cy, 8 * cx), np.float32)
        assert L.ref_transform_to_pixels(t, Kf.ctypes.data, Kf.size, px.ctypes.data, 8 * cx) == 0
        if t in (1, 2, 3, 12, 13, 14, 15, 16, 17):
            dcs = np.array([[K[0, 0]]])
            mine = vd.idct_block(t, K, dcs, 0, 0)
        else:
            mine = vd.Cm(8 * cy) @ (K.T if cy >= cx else K) @ vd.Cm(8 * cx).T
        assert np.abs(mine - px).max() < 2e-5 * max(1.0, np.abs(px).max()), t
