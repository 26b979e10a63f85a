"""Pins the api_level < 34 colour pass (csrc/color_matrix.cc: applyColorMatrix as driven by decodeSampledImageImpl,
ColorMatrix.cpp:35-119 / JniDecoding.cpp:138-228) against the reference run with the Android API level set to 33.
Lossless inputs, so the decode is exact and the pass is what is compared."""
import numpy as np
import pytest

import cases
import hostemu_lib as H

# (name, primaries, transfer) -- jxl/color_encoding.h enums: primaries 1 sRGB, 9 Rec.2100, 11 P3; transfer 13 sRGB, 1 709, 17 DCI, 8 linear
ENCODINGS = [("srgb", 1, 13), ("p3_srgb", 11, 13), ("bt2020_709", 9, 1), ("srgb_709", 1, 1), ("p3_dci", 11, 17), ("srgb_linear", 1, 8),
             ("bt2020_pq", 9, 16), ("bt2020_hlg", 9, 18), ("p3_pq", 11, 16)]


def _source(seed=0, w=96, h=80):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    img[:8] = np.arange(w)[None, :, None] * 255 // (w - 1)   # grey ramp incl. the darkest levels (and a black first pixel)
    # black pixels in the middle of some rows: the reference's tone mapper stops tone-mapping a row at the first one
    img[20, 40] = 0
    img[33, 5] = 0
    img[33, 70] = 0
    img[50, w - 1] = 0
    return img


def encoded(ref, name, prim, tf):
    img = _source()
    h, w, _ = img.shape
    return img, cases._cached("cm2_%s" % name, lambda: ref.encode_ex(img.reshape(-1), w, h, 3, lossless=True, primaries=prim, transfer=tf))


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
    if tf in (16, 18):
        # tone-mapped: one division and three products more per pixel, evaluated by the reference under -ffast-math
        assert d.max() <= 1 and (d != 0).mean() < 2e-3, (d.max(), (d != 0).mean())
        # the row bug is reproduced: a row that starts with a black pixel is converted WITHOUT tone mapping, which
        # differs visibly from a tone-mapped row
        assert (got[0] == want[0]).mean() > 0.998
    else:
        assert d.max() <= 1 and (d != 0).mean() < 1e-4, (d.max(), (d != 0).mean())
    if prim == 1:
        assert d.max() == 0
    if name != "srgb":
        assert (want != r34).mean() > 0.3  # the pass really converts


@pytest.mark.parametrize("enc", [("p3_srgb", 11, 13), ("bt2020_pq", 9, 16), ("bt2020_hlg", 9, 18)], ids=["p3_srgb", "bt2020_pq", "bt2020_hlg"])
def test_color_pass_16bit_matches_reference_api33(enc, ref):
    """applyColorMatrix16Bit (ColorMatrix.cpp:121-219): 16-bit lossless source, api level 33, RGBA_F16 output."""
    name, prim, tf = enc
    rng = np.random.default_rng(3)
    h, w = 64, 80
    img = rng.integers(0, 65536, (h, w, 3)).astype(np.uint16)
    img[:4] = (np.arange(w)[None, :, None] * 65535 // (w - 1)).astype(np.uint16)
    img[20, 30] = 0
    data = cases._cached("cm16_%s" % name, lambda: ref.encode_ex(img.reshape(-1), w, h, 3, bits=16, lossless=True, primaries=prim, transfer=tf))
    raw, meta = ref.decode_oneshot(data)
    assert raw.dtype == np.uint16 and (raw[..., :3] == img).all()
    want = ref.decode_sampled(data, cfg=3, api_level=33)
    wantf = np.ascontiguousarray(want["pixels"][:, : w * 8]).view(np.float16).reshape(h, w, 4)
    st, got = H.color_matrix16(data, raw)
    assert st == 0
    gotf = (got.astype(np.float32) * np.float32(1.0 / 65535.0)).astype(np.float16)   # RgbaU16ToF (imagebit/RgbaU16toHF.cpp:42-144)
    d = np.abs(gotf.astype(np.float32) - wantf.astype(np.float32))
    assert d.max() <= 2.0 / 1024 and (d != 0).mean() < 2e-3, (float(d.max()), float((d != 0).mean()))
