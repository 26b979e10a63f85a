"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path through the C ABI against the committed golden
vectors and, where oracle/_ref is present, against the reference run on the same inputs."""
import numpy as np
import pytest

import cases
import golden_lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def J():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import jxl_coder_b200 as J
    J.load_library()
    return J


@pytest.mark.parametrize("name", golden_lib.names())
def test_decode_matches_golden(J, name):
    g = golden_lib.load(name)
    before = J.kernel_launches()
    bmp = J.JxlCoder.decode(g["jxl"], J.PreferredColorConfig.RGBA_8888)
    assert J.kernel_launches() > before, "no CUDA kernel was launched"
    out = bmp.as_array()
    want = g["raw"].copy()
    if want[..., 3].min() < 255:  # the reference premultiplies (ReformatBitmap.cpp:65-77)
        a = want[..., 3:4].astype(np.uint16)
        want[..., :3] = (want[..., :3].astype(np.uint16) * a // 255).astype(np.uint8)
    if "lossless" in name:
        assert (out == want).all()
    else:
        golden_lib.lossy_close(out, want, name)


@pytest.mark.parametrize("name", ["rgba_lossless_128", "rgb_lossless_200x150"])
@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_color_configs_bit_exact_on_lossless(J, name, cfg):
    g = golden_lib.load(name)
    bmp = J.JxlCoder.decode(g["jxl"], cfg)
    want = g["out_%d" % cfg]
    assert bmp.config == bytes(g["cfg_%d" % cfg]).decode()
    assert bmp.pixels.shape == want.shape
    assert (bmp.pixels == want).all()


@pytest.mark.parametrize("name", cases.SMALL)
def test_decode_matches_reference(J, name, ref):
    data = cases.get(name)
    want = ref.decode_sampled(data, cfg=2)["pixels"].reshape(-1)
    bmp = J.JxlCoder.decode(data, J.PreferredColorConfig.RGBA_8888)
    out = bmp.pixels.reshape(-1)
    assert out.shape == want.shape
    if "lossless" in name:
        assert (out == want).all()
    else:
        golden_lib.lossy_close(out, want, name)


def test_gpu_matches_cpu_emulation_of_same_code(J):
    """The kernels call the same host/device functions tests/hostemu runs serially: results may differ only by
    floating-point contraction (FMA)."""
    import hostemu_lib as H
    for name in ("rgb_lossy_256x200", "natural_d1_e7", "rgba_lossy_300x203"):
        g = golden_lib.load(name)
        e = H.Decoded(g["jxl"])
        emu = e.render()
        e.close()
        if emu[..., 3].min() < 255:
            a = emu[..., 3:4].astype(np.uint16)
            emu[..., :3] = (emu[..., :3].astype(np.uint16) * a // 255).astype(np.uint8)
        out = J.JxlCoder.decode(g["jxl"], 2).as_array()
        d = np.abs(out.astype(int) - emu.astype(int))
        assert d.max() <= 1 and (d == 0).mean() > 0.999, (name, d.max(), (d == 0).mean())


def test_batch_and_errors(J):
    g1 = golden_lib.load("rgb_lossy_64")
    g2 = golden_lib.load("rgba_lossless_128")
    bad = g1["jxl"][:200]
    res = J.decode_batch([g1["jxl"], bad, g2["jxl"], b"garbage"], config=2, raise_on_error=False)
    assert isinstance(res[0], J.Bitmap) and isinstance(res[2], J.Bitmap)
    assert isinstance(res[1], J.InvalidJXLException)
    assert isinstance(res[3], J.InvalidJXLException)
    assert (res[2].as_array()[..., 3] == g2["raw"][..., 3]).all()
    with pytest.raises(J.InvalidJXLException):
        J.JxlCoder.decode(bad)


def test_larger_multi_group_image(J, ref):
    from oracle import synth
    img = synth.synth_image(1024, 768, 11)
    data = cases._cached("rgb_lossy_1024x768", lambda: ref.encode(img, 1024, 768))
    want = ref.decode_sampled(data, cfg=2)["pixels"].reshape(768, 1024, 4)
    out = J.JxlCoder.decode(data, 2).as_array()
    golden_lib.lossy_close(out, want, "1024x768")
