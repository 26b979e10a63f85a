"""Pins the api_level < 34 colour pass (csrc/color_matrix.cc: applyColorMatrix as driven by decodeSampledImageImpl,
ColorMatrix.cpp:35-119 / JniDecoding.cpp:138-228) against the reference run with the Android API level set to 33.
Lossless inputs, so the decode is exact and the pass is what is compared."""
import numpy as np
import pytest

import cases
import hostemu_lib as H

# (name, primaries, transfer) -- jxl/color_encoding.h enums: primaries 1 sRGB, 9 Rec.2100, 11 P3; transfer 13 sRGB, 1 709, 17 DCI, 8 linear
ENCODINGS = [("srgb", 1, 13), ("p3_srgb", 11, 13), ("bt2020_709", 9, 1), ("srgb_709", 1, 1), ("p3_dci", 11, 17), ("srgb_linear", 1, 8)]


def _source(seed=0, w=96, h=80):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    img[:8] = np.arange(w)[None, :, None] * 255 // (w - 1)   # grey ramp incl. the darkest levels
    return img


def encoded(ref, name, prim, tf):
    img = _source()
    h, w, _ = img.shape
    return img, cases._cached("cm_%s" % name, lambda: ref.encode_ex(img.reshape(-1), w, h, 3, lossless=True, primaries=prim, transfer=tf))


@pytest.mark.parametrize("enc", ENCODINGS, ids=[e[0] for e in ENCODINGS])
def test_color_pass_matches_reference_api33(enc, ref):
    name, prim, tf = enc
    img, data = encoded(ref, name, prim, tf)
    h, w, _ = img.shape
    r34 = ref.decode_sampled(data, cfg=2, api_level=34)["pixels"][:, : w * 4].reshape(h, w, 4)
    if name != "srgb_linear":
        assert (r34[..., :3] == img).all()  # lossless: api 34 hands back the stored samples
    want = ref.decode_sampled(data, cfg=2, api_level=33)["pixels"][:, : w * 4].reshape(h, w, 4)
    st, got = H.color_matrix(data, r34)
    if name == "srgb_linear":
        assert st == 1 and (want == r34).all()  # linear transfer: the reference skips the pass
        return
    assert st == 0
    d = np.abs(got.astype(int) - want.astype(int))
    # integer LUT path: exact up to float rounding at a 1/2048 bucket edge of the matrix output
    assert d.max() <= 1 and (d != 0).mean() < 1e-4, (d.max(), (d != 0).mean())
    if prim == 1:
        assert d.max() == 0
    if name != "srgb":
        assert (want != r34).mean() > 0.3  # the pass really converts


def test_pq_and_hlg_are_refused(ref):
    img = _source()
    h, w, _ = img.shape
    for tf in (16, 18):
        data = cases._cached("cm_tf%d" % tf, lambda: ref.encode_ex(img.reshape(-1), w, h, 3, lossless=True, primaries=9, transfer=tf))
        st, _ = H.color_matrix(data, np.zeros((h, w, 4), np.uint8))
        assert st == 2
