"""Synthetic single blocks through the CUDA reconstruction kernels (jxlb_test_recon_block): every plain DCT strategy, in
particular DCT128 / DCT256 / their rectangular variants, which no encoder-made file contains.  Blocks up to 64x64 run in
ReconRegionTmaKernel, larger ones in ReconLargeListKernel.  Checked against libjxl's own TransformToPixels in the reference's
binary, fed with the dequantised coefficient arrays (integer -> float stage checked exactly against the host build of the
same headers, whose own output the CPU suite pins to the reference: tests/test_recon_block_host.py)."""
import ctypes as C

import numpy as np
import pytest

import synth_block as S
from test_recon_block_host import host_recon, reference_pixels

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import jxl_coder_b200 as J
    lib = J.load_library()
    lib.jxlb_test_recon_block.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.jxlb_test_recon_block.restype = C.c_int
    return lib


@pytest.mark.parametrize("dense", [False, True], ids=["sparse", "dense"])
@pytest.mark.parametrize("strategy", S.PLAIN_DCT, ids=[S.NAMES[s] for s in S.PLAIN_DCT])
def test_gpu_block_reconstruction_matches_reference_transform(L, ref, strategy, dense):
    q, lf = S.make(strategy, 2 + int(dense), dense=dense)
    out = np.zeros(q.shape, np.float32)
    assert L.jxlb_test_recon_block(-1, strategy, q.ctypes.data, lf.ctypes.data, 1, 4096, out.ctypes.data) == 0
    co, host_px = host_recon(strategy, q, lf)
    want = reference_pixels(ref, strategy, co)
    peak = float(np.abs(want).max())
    assert peak > 0.05
    assert np.abs(out - want).max() <= 3e-5 * peak, (S.NAMES[strategy], float(np.abs(out - want).max()), peak)
    assert np.abs(out - host_px).max() <= 3e-5 * peak
