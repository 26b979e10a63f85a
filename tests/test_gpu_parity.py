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


def _multi_group_inputs(ref, n=6):
    """n different multi-group lossy images (LF-group kernel, lane-parallel AC, region IDCT, fast filter path)."""
    from oracle import synth
    out = []
    for i in range(n):
        w, h = 512 + 64 * (i % 3), 384 + 40 * (i % 2)
        img = synth.synth_image(w, h, 20 + i)
        out.append(cases._cached("rgb_lossy_%dx%d_s%d" % (w, h, 20 + i), lambda: ref.encode(img, w, h)))
    return out


def test_prepared_batches_overlap_and_match(J, ref):
    """Two prepared batches run asynchronously on their own streams (the benchmark's device-resident mode): every run
    of every batch must reproduce the synchronous result bit for bit, and the span timer must cover the runs."""
    datas = _multi_group_inputs(ref)
    want = [b.as_array().copy() for b in J.decode_batch(datas, config=2)]
    for w, d in zip(want, datas):
        r = ref.decode_sampled(d, cfg=2)
        golden_lib.lossy_close(w, r["pixels"].reshape(w.shape), "multi-group")
    a = J.PreparedBatch(datas, config=2)
    b = J.PreparedBatch(list(reversed(datas)), config=2)
    assert set(a.status) == {0} and set(b.status) == {0}
    a.reset_stats()
    b.reset_stats()
    for _ in range(3):
        assert a.run_async() == 0
        assert b.run_async() == 0
    assert a.wait() == 0 and b.wait() == 0
    for i in range(len(datas)):
        assert (a.fetch(i).as_array() == want[i]).all()
        assert (b.fetch(i).as_array() == want[len(datas) - 1 - i]).all()
    ms, runs = a.stage_ms_mean()
    assert runs == 3 and ms["all_kernels"] > 0 and ms["inverse_transforms"] > 0
    assert a.span_ms(b) > 0 and a.span_ms() > 0
    a.free()
    b.free()


def test_concurrent_callers_match(J, ref):
    """jxlb_decode_batch is re-entrant: four threads decoding at once (separate decode slots on one GPU) return the
    same pixels as a single caller."""
    import threading
    datas = _multi_group_inputs(ref)
    want = [b.as_array().copy() for b in J.decode_batch(datas, config=2)]
    errs = []

    def work(k):
        try:
            for _ in range(2):
                order = list(range(len(datas)))
                order = order[k:] + order[:k]
                res = J.decode_batch([datas[i] for i in order], config=2)
                for i, r in zip(order, res):
                    if not (r.as_array() == want[i]).all():
                        errs.append((k, i))
        except Exception as e:  # noqa
            errs.append(repr(e))
    ths = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs


@pytest.mark.parametrize("cfg", [3, 5])
def test_f16_and_1010102_on_lossy(J, ref, cfg):
    """configs[2] / configs[3] output formats on a lossy multi-group image: RGBA_F16 and RGBA_1010102 are integer
    re-packings of the 8-bit decode (Rgba8ToF16 / Rgba8ToRGBA1010102), so they must equal the reference's packing of
    OUR 8-bit pixels exactly, and the reference's own output within the lossy tolerance."""
    data = _multi_group_inputs(ref, 1)[0]
    got = J.JxlCoder.decode(data, cfg)
    r = ref.decode_sampled(data, cfg=cfg)
    assert got.width == r["width"] and got.height == r["height"]
    bpp = 8 if cfg == 3 else 4
    mine = np.ascontiguousarray(got.pixels[:, : got.width * bpp])
    theirs = np.ascontiguousarray(r["pixels"][:, : got.width * bpp])
    if cfg == 3:
        a = mine.view(np.float16).astype(np.float32)
        b = theirs.view(np.float16).astype(np.float32)
        assert np.abs(a - b).max() <= 1.0 / 255 + 1e-3
        assert (a == b).mean() > 0.95
    else:
        a = mine.view(np.uint32)
        b = theirs.view(np.uint32)
        for sh in (0, 10, 20):
            d = np.abs(((a >> sh) & 0x3FF).astype(int) - ((b >> sh) & 0x3FF).astype(int))
            assert d.max() <= 4 and (d == 0).mean() > 0.95  # 1 LSB of 8 bits = 4 of 10
        assert ((a >> 30) == (b >> 30)).all()


@pytest.mark.parametrize("kind", ["rgb_lossy", "rgba_lossless", "rgba_lossy"])
def test_animated_frames_match_reference(J, ref, kind):
    """JxlAnimatedImage.getFrame(i) on animations written by the reference's own JxlAnimatedEncoder (full-canvas kReplace
    frames, configs[4] shape): frame table identical; lossy RGB frames within the lossy tolerance of the reference's
    coalesced frames (lossy RGBA -- configs[4]'s content, alpha coded with the squeeze transform -- with a bit-exact alpha
    plane), lossless RGBA frames bit-exact."""
    data = cases.anim_case(kind)
    w, h, n = cases.ANIM_W, cases.ANIM_H, cases.ANIM_N
    ra = ref.Anim(data, cfg=2)
    a = J.JxlAnimatedImage(data, J.PreferredColorConfig.RGBA_8888)
    assert a.number_of_frames == len(ra) == n
    assert (a.get_width(), a.get_height()) == ra.size == (w, h)
    for i in list(range(n)) + [1, 3, 0]:  # sequential access is served from the prefetched batch, out-of-order access re-decodes
        assert a.get_frame_duration(i) == ra.duration(i)
        want = ra.frame(i)["pixels"][:, : w * 4].reshape(h, w, 4)
        got = a.get_frame(i).as_array()
        if kind == "rgba_lossless":
            assert (got == want).all()
        elif kind == "rgba_lossy":
            assert (got[..., 3] == want[..., 3]).all()
            golden_lib.lossy_close(got, want, "anim frame %d" % i)
        else:
            golden_lib.lossy_close(got, want, "anim frame %d" % i)
    with pytest.raises(J.JxlCoderError):
        a.get_frame(n)
    a.close()
    ra.close()


# ---- decodeSampled with a target size (RescaleImage -> weave_scale_u8) ----
def _resize_src(ref, w, h):
    import test_resize_host as T
    img = T._image(w, h, w * 1000 + h)
    data = cases._cached("resize_src_%dx%d" % (w, h), lambda: ref.encode(img[..., :3].reshape(-1), w, h, colorspace=1, compression=1))
    return img, data


@pytest.mark.parametrize("case", __import__("test_resize_host").CASES)
def test_decode_sampled_rescale_bit_exact_on_lossless(J, ref, case):
    """Lossless sources decode exactly, so the rescaled picture must equal the reference's bit for bit."""
    w, h, rw, rh, mode, filt = case
    _, data = _resize_src(ref, w, h)
    r = ref.decode_sampled(data, w=rw, h=rh, cfg=2, scale_mode=mode, filt=filt)
    want = r["pixels"][:, : r["width"] * 4].reshape(r["height"], r["width"], 4)
    before = J.kernel_launches()
    got = J.JxlCoder.decode_sampled(data, rw, rh, 2, mode, filt)
    assert J.kernel_launches() > before
    assert (got.width, got.height) == (r["width"], r["height"])
    import hostemu_lib as H
    import test_resize_host as T
    T._check(got.as_array(), want, filt, case)  # the reference within the bound pinned in test_resize_host.py
    img, _ = _resize_src(ref, w, h)
    assert (got.as_array() == H.resize_rgba8(img, rw, rh, mode, filt)).all()  # the kernels == their CPU restatement, always


@pytest.mark.parametrize("cfg", [3, 4, 5])
def test_decode_sampled_rescale_then_reformat(J, ref, cfg):
    """ReformatColorConfig runs on the rescaled picture (RGBA_F16, RGB_565, RGBA_1010102 -- configs[3]'s output format)."""
    _, data = _resize_src(ref, 200, 120)
    r = ref.decode_sampled(data, w=50, h=30, cfg=cfg, scale_mode=1, filt=4)
    got = J.JxlCoder.decode_sampled(data, 50, 30, cfg, 1, 4)
    assert (got.width, got.height) == (r["width"], r["height"])
    bpp = {3: 8, 4: 2, 5: 4}[cfg]
    assert (got.pixels[:, : got.width * bpp] == r["pixels"][:, : got.width * bpp]).all()


def test_decode_sampled_rescale_on_lossy_multi_group(J, ref):
    """configs[3] in small: lossy VarDCT multi-group image -> FIT + Mitchell -> RGBA_1010102.  The rescaler must be
    bit-exact on OUR decoded pixels (CPU restatement pinned against the reference in test_resize_host.py), and the
    whole call within the lossy tolerance of the reference."""
    import hostemu_lib as H
    from oracle import synth
    w, h = 1024, 768
    img = synth.synth_image(w, h, 11)
    data = cases._cached("rgb_lossy_1024x768", lambda: ref.encode(img, w, h))
    full = J.JxlCoder.decode(data, 2).as_array()
    for (rw, rh, mode, filt) in [(256, 192, 1, 4), (300, 300, 1, 6), (333, 77, 3, 1), (512, -1, 1, 7)]:
        got = J.JxlCoder.decode_sampled(data, rw, rh, 2, mode, filt).as_array()
        mine = H.resize_rgba8(full, rw, rh, mode, filt)
        assert got.shape == mine.shape and (got == mine).all(), (rw, rh, mode, filt)
        r = ref.decode_sampled(data, w=rw, h=rh, cfg=2, scale_mode=mode, filt=filt)
        want = r["pixels"][:, : r["width"] * 4].reshape(r["height"], r["width"], 4)
        d = np.abs(got.astype(int) - want.astype(int))
        assert d.max() <= 1 and (d == 0).mean() > 0.97, (d.max(), (d == 0).mean())
    got = J.JxlCoder.decode_sampled(data, 256, 192, 5, 1, 4)
    r = ref.decode_sampled(data, w=256, h=192, cfg=5, scale_mode=1, filt=4)
    a = np.ascontiguousarray(got.pixels[:, : 256 * 4]).view(np.uint32)
    b = np.ascontiguousarray(r["pixels"][:, : 256 * 4]).view(np.uint32)
    for sh in (0, 10, 20):
        d = np.abs(((a >> sh) & 0x3FF).astype(int) - ((b >> sh) & 0x3FF).astype(int))
        assert d.max() <= 4 and (d == 0).mean() > 0.97
    assert ((a >> 30) == (b >> 30)).all()


def test_decode_sampled_lanczos_hann_bicubic(J, ref):
    """Lanczos3, HANN (mapped to Lanczos3 by SizeScaler.cpp:86-89) and Bicubic on a lossless source: bit-exact.  Lanczos3 on a
    source with alpha is refused (resize.h)."""
    _, data = _resize_src(ref, 96, 64)
    for (rw, rh, mode, filt) in [(24, 16, 3, 5), (24, 16, 3, 9), (40, 40, 2, 10), (150, 100, 1, 5), (35, 23, 3, 10)]:
        r = ref.decode_sampled(data, w=rw, h=rh, cfg=2, scale_mode=mode, filt=filt)
        got = J.JxlCoder.decode_sampled(data, rw, rh, 2, mode, filt)
        assert (got.width, got.height) == (r["width"], r["height"])
        d = np.abs(got.pixels[:, : got.width * 4].astype(int) - r["pixels"][:, : got.width * 4].astype(int))
        if filt == 10:   # the spline family: one Q15 tap weight in ~10^3 off by one unit on inexact ratios (resize.h)
            assert d.max() <= 1 and (d != 0).mean() < 2e-3, (rw, rh, mode, filt, d.max(), (d != 0).mean())
        else:
            assert d.max() == 0, (rw, rh, mode, filt)
    with pytest.raises(J.UnsupportedJXLException):
        J.JxlCoder.decode_sampled(cases.get("rgba_lossless_128"), 40, 40, 2, 1, 5)
    # w = h = -1: no rescale (JxlCoder.kt:55-62); 0 on an axis: no rescale either (JniDecoding.cpp:116-117)
    assert J.JxlCoder.decode_sampled(data, -1, -1, 2, 1, 4).width == 96
    assert J.JxlCoder.decode_sampled(data, 0, 10, 2, 1, 4).width == 96


def test_decode_sampled_rescale_with_alpha(J, ref):
    """RGBA lossless source: premultiply / rescale / divide back, then ReformatColorConfig premultiplies the result."""
    import hostemu_lib as H
    import test_resize_host as T
    from oracle import synth
    img = synth.synth_image(128, 128, 3, alpha=True).reshape(128, 128, 4)
    data = cases.get("rgba_lossless_128")
    for (rw, rh, mode, filt) in [(40, 40, 1, 4), (64, 20, 2, 1), (200, 150, 3, 6), (33, 77, 3, 2), (128, 50, 3, 7)]:
        r = ref.decode_sampled(data, w=rw, h=rh, cfg=2, scale_mode=mode, filt=filt)
        want = r["pixels"][:, : r["width"] * 4].reshape(r["height"], r["width"], 4)
        got = J.JxlCoder.decode_sampled(data, rw, rh, 2, mode, filt).as_array()
        T._check(got, want, filt, (rw, rh, mode, filt))
        mine = H.resize_rgba8(img, rw, rh, mode, filt, has_alpha=True)
        a = mine[..., 3:4].astype(np.uint16)
        mine[..., :3] = (mine[..., :3].astype(np.uint16) * a // 255).astype(np.uint8)
        assert (got == mine).all(), (rw, rh, mode, filt)


def test_animated_frame_rescale(J, ref):
    data = cases.anim_case("rgb_lossy")
    ra = ref.Anim(data, cfg=2, scale_mode=1, filt=4)
    a = J.JxlAnimatedImage(data, J.PreferredColorConfig.RGBA_8888, J.ScaleMode.FIT, J.JxlResizeFilter.MITCHELL_NETRAVALI)
    r = ra.frame(2, 80, 66)
    want = r["pixels"][:, : r["width"] * 4].reshape(r["height"], r["width"], 4)
    got = a.get_frame(2, 80, 66).as_array()
    assert got.shape == want.shape
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 1 and (d == 0).mean() > 0.97
    a.close()
    ra.close()


def test_config3_full_size_8k_to_1080p_1010102(J, ref):
    """BASELINE configs[3] at full size: 7680x4320 lossy -> decodeSampled(1920, 1080, RGBA_1010102, FIT, Mitchell).
    Against the reference within the lossy tolerance (1 LSB of 8 bits = 4 of 10), and the rescale bit-exact on our own
    decoded pixels (CPU restatement of the rescaler)."""
    import hostemu_lib as H
    from oracle import gen_inputs
    data = gen_inputs.c4_image()
    got = J.JxlCoder.decode_sampled(data, 1920, 1080, 5, 1, 4)
    assert (got.width, got.height) == (1920, 1080)
    a = np.ascontiguousarray(got.pixels[:, : 1920 * 4]).view(np.uint32)
    full = J.JxlCoder.decode(data, 2).as_array()
    mine = H.resize_rgba8(full, 1920, 1080, 1, 4).astype(np.uint32)
    packed = (mine[..., 0] << 2) | (mine[..., 1] << 12) | (mine[..., 2] << 22) | ((mine[..., 3] >> 6) << 30)
    assert (a == packed).all()
    r = ref.decode_sampled(data, w=1920, h=1080, cfg=5, scale_mode=1, filt=4)
    b = np.ascontiguousarray(r["pixels"][:, : 1920 * 4]).view(np.uint32)
    for sh in (0, 10, 20):
        d = np.abs(((a >> sh) & 0x3FF).astype(int) - ((b >> sh) & 0x3FF).astype(int))
        assert d.max() <= 4 and (d == 0).mean() > 0.97, (sh, d.max(), (d == 0).mean())
    assert ((a >> 30) == (b >> 30)).all()


def test_config2_full_size_1080p_f16(J, ref):
    """BASELINE configs[2] at full size (one image of the batch): 1920x1080 lossy -> RGBA_F16."""
    from oracle import gen_inputs
    data = gen_inputs.c3_image(0)
    got = J.JxlCoder.decode(data, 3)
    r = ref.decode_sampled(data, cfg=3)
    a = np.ascontiguousarray(got.pixels[:, : 1920 * 8]).view(np.float16).astype(np.float32)
    b = np.ascontiguousarray(r["pixels"][:, : 1920 * 8]).view(np.float16).astype(np.float32)
    assert np.abs(a - b).max() <= 2.0 / 255 + 1e-3 and (a == b).mean() > 0.97  # one 8-bit step of the decode, then the resampler


# ---- api_level < 34 colour pass (applyColorMatrix) ----
@pytest.mark.parametrize("enc", __import__("test_color_matrix_host").ENCODINGS, ids=[e[0] for e in __import__("test_color_matrix_host").ENCODINGS])
def test_api33_color_pass_matches_reference(J, ref, enc):
    import test_color_matrix_host as T
    name, prim, tf = enc
    img, data = T.encoded(ref, name, prim, tf)
    h, w, _ = img.shape
    if name == "srgb_linear":  # the reference converts these through lcms2 (ICC path): refused, not handed back unconverted
        with pytest.raises(J.UnsupportedJXLException):
            J.JxlCoder.decode(data, 2)
        return
    want = ref.decode_sampled(data, cfg=2, api_level=33)
    old = J.JxlCoder.api_level
    J.JxlCoder.api_level = 33
    try:
        got = J.JxlCoder.decode(data, 2)
        assert got.color_space == ""  # no ColorSpace tag below API 34 (JniDecoding.cpp:236)
        d = np.abs(got.as_array().astype(int) - want["pixels"][:, : w * 4].reshape(h, w, 4).astype(int))
        assert d.max() <= 1 and (d != 0).mean() < (2e-3 if tf in (16, 18) else 1e-4)   # PQ / HLG: tone-mapped (Rec2408ToneMapper)
        if prim == 1:
            assert d.max() == 0
        # after the rescale, before the reformat (1010102)
        r = ref.decode_sampled(data, w=40, h=30, cfg=2, scale_mode=3, filt=1, api_level=33)
        g2 = J.JxlCoder.decode_sampled(data, 40, 30, 2, 3, 1).as_array()
        d = np.abs(g2.astype(int) - r["pixels"][:, : 40 * 4].reshape(30, 40, 4).astype(int))
        assert d.max() <= 1 and (d != 0).mean() < 1e-3
    finally:
        J.JxlCoder.api_level = old


def test_api33_on_lossy_and_refusals(J, ref):
    data = cases.get("rgb_lossy_256x200")
    want = ref.decode_sampled(data, cfg=2, api_level=33)["pixels"][:, : 256 * 4].reshape(200, 256, 4)
    old = J.JxlCoder.api_level
    J.JxlCoder.api_level = 33
    try:
        got = J.JxlCoder.decode(data, 2).as_array()
        d = np.abs(got.astype(int) - want.astype(int))
        assert d.max() <= 2 and (d == 0).mean() > 0.97  # 1 LSB of the lossy decode through the 2048-level requantisation
    finally:
        J.JxlCoder.api_level = old


@pytest.mark.parametrize("enc", [("p3_srgb", 11, 13), ("bt2020_pq", 9, 16), ("bt2020_hlg", 9, 18)], ids=["p3_srgb", "bt2020_pq", "bt2020_hlg"])
def test_api33_color_pass_16bit(J, ref, enc):
    """applyColorMatrix16Bit through the GPU path: 16-bit lossless source, api level 33, RGBA_F16 and RGBA_8888 output."""
    name, prim, tf = enc
    rng = np.random.default_rng(3)
    h, w = 64, 80
    img = rng.integers(0, 65536, (h, w, 3)).astype(np.uint16)
    img[:4] = (np.arange(w)[None, :, None] * 65535 // (w - 1)).astype(np.uint16)
    img[20, 30] = 0
    data = cases._cached("cm16_%s" % name, lambda: ref.encode_ex(img.reshape(-1), w, h, 3, bits=16, lossless=True, primaries=prim, transfer=tf))
    old = J.JxlCoder.api_level
    J.JxlCoder.api_level = 33
    try:
        want = ref.decode_sampled(data, cfg=3, api_level=33)
        got = J.JxlCoder.decode(data, 3)
        a = np.ascontiguousarray(got.pixels[:, : w * 8]).view(np.float16).astype(np.float32)
        b = np.ascontiguousarray(want["pixels"][:, : w * 8]).view(np.float16).astype(np.float32)
        d = np.abs(a - b)
        assert d.max() <= 2.0 / 1024 and (d != 0).mean() < 2e-3, (float(d.max()), float((d != 0).mean()))
        want8 = ref.decode_sampled(data, cfg=2, api_level=33)["pixels"][:, : w * 4]
        got8 = J.JxlCoder.decode(data, 2).pixels[:, : w * 4]
        d8 = np.abs(got8.astype(int) - want8.astype(int))
        assert d8.max() <= 1 and (d8 != 0).mean() < 2e-3
    finally:
        J.JxlCoder.api_level = old


# ---- codestream orientation ----
@pytest.mark.parametrize("o", range(2, 9))
def test_orientation_matches_reference(J, ref, o):
    import test_orientation_host as T
    for alpha in (False, True):
        img, data = T.encoded(ref, o, alpha)
        r = ref.decode_sampled(data, cfg=2)
        want = r["pixels"][:, : r["width"] * 4].reshape(r["height"], r["width"], 4)
        got = J.JxlCoder.decode(data, 2)
        assert (got.width, got.height) == (r["width"], r["height"]) == J.JxlCoder.get_size(data)
        assert (got.as_array() == want).all()
    # orientation, then rescale, then reformat
    img, data = T.encoded(ref, o, False)
    r = ref.decode_sampled(data, w=20, h=30, cfg=5, scale_mode=3, filt=1)
    got = J.JxlCoder.decode_sampled(data, 20, 30, 5, 3, 1)
    assert (got.pixels[:, : 20 * 4] == r["pixels"][:, : 20 * 4]).all()


@pytest.mark.parametrize("o", range(2, 9))
def test_orientation_on_lossy(J, ref, o):
    from oracle import synth
    w, h = 200, 136
    img = synth.synth_image(w, h, 5)
    data = cases._cached("orient%d_lossy" % o, lambda: ref.encode_ex(img, w, h, 3, distance=1.0, orientation=o))
    r = ref.decode_sampled(data, cfg=2)
    got = J.JxlCoder.decode(data, 2)
    assert (got.width, got.height) == (r["width"], r["height"])
    golden_lib.lossy_close(got.as_array(), r["pixels"][:, : r["width"] * 4].reshape(r["height"], r["width"], 4), "orientation %d lossy" % o)


# ---- the 13 JXL files the reference's demo app ships (app/src/main/assets; copied into tests/_cache/assets, not committed) ----
def _assets():
    import glob
    import os
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_cache", "assets")
    return sorted(glob.glob(os.path.join(d, "*.jxl")))


@pytest.mark.parametrize("path", _assets(), ids=[__import__("os").path.basename(p) for p in _assets()])
def test_reference_app_assets(J, ref, path):
    """Every file either decodes to the reference's pixels (lossless exact / lossy bound) or is refused with
    UnsupportedJXLException -- never a different picture."""
    data = open(path, "rb").read()
    r = ref.decode_sampled(data, cfg=2)
    want = r["pixels"][:, : r["width"] * 4].reshape(r["height"], r["width"], 4)
    assert J.JxlCoder.get_size(data) == (r["width"], r["height"])
    try:
        got = J.JxlCoder.decode(data, 2)
    except J.UnsupportedJXLException as e:
        pytest.skip("refused: %s" % e)
    assert (got.width, got.height) == (r["width"], r["height"])
    d = np.abs(got.as_array().astype(int) - want.astype(int))
    assert d.max() <= 1 and (d == 0).mean() > 0.97, (d.max(), (d == 0).mean())


# ---- lossy alpha: squeeze transform ----
@pytest.mark.parametrize("shape", [(120, 90, 40, 1.0), (320, 264, 41, 1.0), (257, 300, 42, 2.0), (700, 520, 43, 0.5), (1024, 300, 44, 1.0), (2304, 300, 45, 1.0), (520, 4200, 46, 1.0)])
def test_squeezed_alpha_matches_reference(J, ref, shape):
    import test_squeeze_host as T
    w, h, seed, ad = shape
    data = T.rgba_lossy(ref, w, h, seed, ad)
    r = ref.decode_sampled(data, cfg=2)
    want = r["pixels"][:, : w * 4].reshape(h, w, 4)
    got = J.JxlCoder.decode(data, 2)
    assert got.premultiplied
    out = got.as_array()
    assert (out[..., 3] == want[..., 3]).all()
    golden_lib.lossy_close(out, want, "squeezed alpha %dx%d" % (w, h))
    # rescale of a source with alpha (premultiply around the passes), then reformat
    r = ref.decode_sampled(data, w=w // 3, h=h // 3, cfg=3, scale_mode=3, filt=1)
    g2 = J.JxlCoder.decode_sampled(data, w // 3, h // 3, 3, 3, 1)
    a = np.ascontiguousarray(g2.pixels[:, : (w // 3) * 8]).view(np.float16).astype(np.float32)
    b = np.ascontiguousarray(r["pixels"][:, : (w // 3) * 8]).view(np.float16).astype(np.float32)
    assert np.abs(a - b).max() <= 2.0 / 255 + 1e-3 and (a == b).mean() > 0.97  # one 8-bit step of the decode, then the resampler


def test_corrupt_inputs_never_hang(J):
    """168 mutated / truncated files through jxlb_decode_batch in a subprocess under a timeout: each comes back as a picture
    or as an error, nothing hangs or crashes, and the library still decodes afterwards."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(here), "tools", "probes", "gpu_fuzz.py"), "7"], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "fuzz ok=" in r.stdout


def test_cropped_frames_over_empty_canvas(J, ref):
    """The reference's animated demo asset: every frame after the first is a cropped kReplace frame whose source slot was
    never written, i.e. the frame's bounding box over a cleared canvas.  All frames against the reference's coalesced ones."""
    import os
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_cache", "assets", "animated_jxl.jxl")
    if not os.path.exists(p):
        pytest.skip("asset not cached")
    data = open(p, "rb").read()
    ra = ref.Anim(data, cfg=2)
    a = J.JxlAnimatedImage(data, J.PreferredColorConfig.RGBA_8888)
    assert a.number_of_frames == len(ra) and (a.get_width(), a.get_height()) == ra.size
    w, h = ra.size
    for i in range(len(ra)):
        want = ra.frame(i)["pixels"][:, : w * 4].reshape(h, w, 4)
        got = a.get_frame(i).as_array()
        assert a.get_frame_duration(i) == ra.duration(i)
        assert (got[..., 3] == want[..., 3]).all(), i
        d = np.abs(got.astype(int) - want.astype(int))
        assert d.max() <= 1 and (d == 0).mean() > 0.97, (i, d.max(), (d == 0).mean())
    a.close()
    ra.close()


@pytest.mark.parametrize("o", [3, 6, 7])
@pytest.mark.parametrize("cfg", [1, 2, 3])
def test_orientation_on_16bit_lossy(J, ref, o, cfg):
    """Sources deeper than 8 bits are staged as RGBA16 in front of the orientation pass (found by tools/probes/gpu_sweep.py)."""
    from oracle import synth
    w, h = 218, 155
    img = synth.synth_image(w, h, 8).astype(np.uint16) * 257
    data = cases._cached("orient%d_lossy16" % o, lambda: ref.encode_ex(img, w, h, 3, bits=16, distance=1.0, orientation=o))
    r = ref.decode_sampled(data, cfg=cfg)
    got = J.JxlCoder.decode(data, cfg)
    assert (got.width, got.height) == (r["width"], r["height"])
    if cfg == 2:
        golden_lib.lossy_close(got.as_array(), r["pixels"][:, : got.width * 4].reshape(got.height, got.width, 4), "16-bit orientation %d" % o)
    elif cfg == 3:
        a = np.ascontiguousarray(got.pixels[:, : got.width * 8]).view(np.float16).astype(np.float32)
        b = np.ascontiguousarray(r["pixels"][:, : got.width * 8]).view(np.float16).astype(np.float32)
        assert np.abs(a - b).max() <= 1.0 / 255
    else:  # DEFAULT -> RGBA_1010102 (depth > 8, no alpha, api >= 33)
        assert got.config == "RGBA_1010102"
        a = np.ascontiguousarray(got.pixels[:, : got.width * 4]).view(np.uint32)
        b = np.ascontiguousarray(r["pixels"][:, : got.width * 4]).view(np.uint32)
        for sh in (0, 10, 20):
            assert np.abs(((a >> sh) & 0x3FF).astype(int) - ((b >> sh) & 0x3FF).astype(int)).max() <= 4


@pytest.mark.parametrize("epf", [1, 2, 3])
def test_epf_active_everywhere(J, ref, epf):
    """Effort-2 encodes filter every cell (sharpness 4): pins the EPF normalisation (libjxl's approximate reciprocal)."""
    from oracle import synth
    w, h = 311, 231
    img = synth.synth_image(w, h, 3)
    data = cases._cached("epf%d_effort2_311x231" % epf, lambda: ref.encode_ex(img, w, h, 3, distance=1.0, options={"EFFORT": 2, "EPF": epf, "GABORISH": 1}))
    want = ref.decode_sampled(data, cfg=2)["pixels"][:, : w * 4].reshape(h, w, 4)
    got = J.JxlCoder.decode(data, 2).as_array()
    d = np.abs(got[..., :3].astype(int) - want[..., :3].astype(int))
    assert d.max() <= 1 and (d == 0).mean() > 0.985, (d.max(), (d == 0).mean())


def test_submit_collect_matches_sync_call(J, ref):
    """jxlb_decode_batch_submit / _collect: three batches in flight from one thread, inputs released right after submit,
    results identical to the synchronous call and within the lossy bound of the reference."""
    datas = [cases.get(n) for n in cases.SMALL[:6]]
    sync = [b.pixels.copy() for b in J.decode_batch(datas, config=2)]
    pend = []
    for k in range(3):
        copies = [bytes(bytearray(d)) for d in datas]
        pend.append(J.PendingBatch(copies, config=2))
        del copies
    for p in pend:
        res = p.result()
        assert len(res) == len(datas)
        for got, want in zip(res, sync):
            assert (got.pixels == want).all()
    # errors stay per image
    p = J.PendingBatch([datas[0], b"not a jxl file", datas[1]], config=2)
    res = p.result(raise_on_error=False)
    assert isinstance(res[1], J.JxlCoderError) and not isinstance(res[0], Exception) and not isinstance(res[2], Exception)
