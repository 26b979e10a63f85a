"""GPU parity of frame composition (what the reference gets from libjxl's coalescing,
interop/JxlAnimatedDecoder.cpp:52-54 and DecodeJpegXlOneShot's last-frame-wins loop): cropped frames, reference slots, the
blend modes kReplace / kAdd / kBlend / kAlphaWeightedAdd / kMul, stills made of layers and animations whose frames are
patches over the previous picture.  Inputs are staged with the reference's own libjxl encoder (oracle/refjxl.encode_layers)."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def J():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import jxl_coder_b200 as J
    J.load_library()
    return J


def _layers(mode, channels, seed=0, w=96, h=80):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h, w, channels)).astype(np.uint8)
    ov = rng.integers(0, 256, (30, 40, channels)).astype(np.uint8)
    if channels == 4:
        base[..., 3] = rng.integers(128, 256, (h, w))
        ov[..., 3] = rng.integers(0, 256, (30, 40))
        ov[:5, :5, 3] = 0
        ov[5:10, :5, 3] = 255
    return [dict(pixels=base, mode=0, source=0, save=1, duration=0), dict(pixels=ov, x0=10, y0=20, mode=mode, source=1, save=0, duration=0)]


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("channels", [4, 3])
def test_two_layer_still_lossless(J, ref, mode, channels):
    w, h = 96, 80
    data = cases._cached("layers_%d_%d" % (mode, channels), lambda: ref.encode_layers(_layers(mode, channels), w, h, channels=channels))
    want = ref.decode_sampled(data, cfg=2)["pixels"][:, : w * 4].reshape(h, w, 4)
    got = J.JxlCoder.decode(data, J.PreferredColorConfig.RGBA_8888).as_array()
    d = np.abs(got.astype(int) - want.astype(int))
    # integer samples blended in f32 like libjxl: exact for kReplace, else at most one unit on a few samples
    if mode == 0:
        assert d.max() == 0
    else:
        assert d.max() <= 1 and (d != 0).mean() < 5e-3, (mode, channels, int(d.max()), float((d != 0).mean()))


def test_layer_sticking_out_of_the_canvas(J, ref):
    w, h = 64, 48
    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, (h, w, 4)).astype(np.uint8)
    ov = rng.integers(0, 256, (40, 50, 4)).astype(np.uint8)
    layers = [dict(pixels=base, mode=0, source=0, save=2, duration=0), dict(pixels=ov, x0=-10, y0=30, mode=2, source=2, save=0, duration=0)]
    data = cases._cached("layers_out", lambda: ref.encode_layers(layers, w, h))
    want = ref.decode_sampled(data, cfg=2)["pixels"][:, : w * 4].reshape(h, w, 4)
    got = J.JxlCoder.decode(data, 2).as_array()
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 5e-3


def _patch_animation(n=6, w=128, h=96, seed=1):
    rng = np.random.default_rng(seed)
    full = rng.integers(0, 256, (h, w, 4)).astype(np.uint8)
    full[..., 3] = 255
    layers = [dict(pixels=full, mode=0, source=0, save=1, duration=40)]
    for i in range(1, n):
        pw, ph = int(rng.integers(16, 64)), int(rng.integers(16, 48))
        patch = rng.integers(0, 256, (ph, pw, 4)).astype(np.uint8)
        patch[..., 3] = rng.integers(0, 256, (ph, pw))
        layers.append(dict(pixels=patch, x0=int(rng.integers(0, w - pw)), y0=int(rng.integers(0, h - ph)), mode=2 if i % 2 else 0, source=1, save=1,
                           duration=40))
    return layers


def test_animation_of_patches_over_the_previous_frame(J, ref):
    w, h = 128, 96
    data = cases._cached("layers_anim", lambda: ref.encode_layers(_patch_animation(), w, h, animation=True))
    a = J.JxlAnimatedImage(data, J.PreferredColorConfig.RGBA_8888)
    r = ref.Anim(data, cfg=2)
    assert a.number_of_frames == len(r) == 6
    for i in [0, 3, 1, 5, 2, 4]:   # any order: every getFrame composes from the start of its dependency chain
        want = r.frame(i)["pixels"][:, : w * 4].reshape(h, w, 4)
        got = a.get_frame(i).as_array()
        d = np.abs(got.astype(int) - want.astype(int))
        assert d.max() <= 1 and (d != 0).mean() < 5e-3, (i, int(d.max()), float((d != 0).mean()))
    a.close()
    r.close()
    # decode() of the file hands back the last frame, composed
    want = ref.decode_sampled(data, cfg=2)["pixels"][:, : w * 4].reshape(h, w, 4)
    got = J.JxlCoder.decode(data, 2).as_array()
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1


def test_lossy_layers_and_rescale(J, ref):
    """Lossy (VarDCT) layers: a cropped kReplace frame over a saved lossy base, then the same through decodeSampled."""
    from oracle import synth
    w, h = 320, 240
    base = synth.synth_image(w, h, 7, alpha=False).reshape(h, w, 3)
    ov = synth.synth_image(100, 80, 8, alpha=False).reshape(80, 100, 3)
    layers = [dict(pixels=base, mode=0, source=0, save=1, duration=0), dict(pixels=ov, x0=50, y0=60, mode=0, source=1, save=0, duration=0)]
    data = cases._cached("layers_lossy_rgb", lambda: ref.encode_layers(layers, w, h, channels=3, lossless=False, distance=1.0, effort=7))
    want = ref.decode_sampled(data, cfg=2)["pixels"][:, : w * 4].reshape(h, w, 4)
    got = J.JxlCoder.decode(data, 2).as_array()
    d = np.abs(got.astype(int) - want.astype(int))
    # the reference keeps unrounded floats until the end; the layers here are rounded (and dithered) when they are decoded
    assert d.max() <= 1 and (d == 0).mean() > 0.97, (int(d.max()), float((d == 0).mean()))
    r2 = ref.decode_sampled(data, w=160, h=120, cfg=2, scale_mode=3, filt=1)
    g2 = J.JxlCoder.decode_sampled(data, 160, 120, 2, 3, 1).as_array()
    d2 = np.abs(g2.astype(int) - r2["pixels"][:, : 160 * 4].reshape(120, 160, 4).astype(int))
    assert d2.max() <= 1 and (d2 == 0).mean() > 0.95
