"""Single synthetic blocks of every plain DCT strategy -- including DCT128 / DCT256 families, which libjxl's encoder never
emits -- through the reconstruction code of csrc/recon.h + numeric.h compiled for the host, against libjxl's own
TransformToPixels in the reference's shipped binary (oracle/ref_api.cpp: ref_transform_to_pixels)."""
import ctypes as C

import numpy as np
import pytest

import hostemu_lib as H
import synth_block as S


def host_recon(strategy, q, lf, hf_mul=1, global_scale=4096):
    L = H.lib()
    L.emu_recon_block.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    co = np.zeros(q.shape, np.float32)
    px = np.zeros(q.shape, np.float32)
    assert L.emu_recon_block(strategy, q.ctypes.data, lf.ctypes.data, hf_mul, global_scale, co.ctypes.data, px.ctypes.data) == 0
    return co, px


def reference_pixels(ref, strategy, coeffs):
    L = ref.lib()
    L.ref_transform_to_pixels.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    out = np.zeros(coeffs.shape, np.float32)
    for c in range(3):
        K = S.libjxl_layout(strategy, coeffs[c])
        px = np.zeros(coeffs[c].shape, np.float32)
        assert L.ref_transform_to_pixels(strategy, K.ctypes.data, K.size, px.ctypes.data, px.shape[1]) == 0
        out[c] = px
    return out


@pytest.mark.parametrize("strategy", S.PLAIN_DCT, ids=[S.NAMES[s] for s in S.PLAIN_DCT])
def test_block_reconstruction_matches_reference_transform(strategy, ref):
    q, lf = S.make(strategy, 1)
    co, px = host_recon(strategy, q, lf)
    want = reference_pixels(ref, strategy, co)
    peak = float(np.abs(want).max())
    assert peak > 0.05  # the block carries signal
    assert np.abs(px - want).max() <= 2e-5 * peak, (S.NAMES[strategy], float(np.abs(px - want).max()), peak)
