"""GPU parity on the BASELINE.json inputs themselves, at full size, against the reference (oracle/_ref) run on the same
files: every distinct 4096x4096 image of configs[1] (not a spot check of image 0), the 512x512 lossless RGBA file of
configs[0] bit-exact, frames of the 120-frame 1024x1024 animation of configs[4], the four distinct 1080p pictures of
configs[2] in RGBA_F16 inside a 256-image batch; the colour-space tag (JniDecoding.cpp:236-253) at API level 34; and
Rec.2100 PQ output."""
import numpy as np
import pytest

import golden_lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def J():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import jxl_coder_b200 as J
    J.load_library()
    return J


@pytest.fixture(scope="module")
def gen(ref):
    from oracle import gen_inputs
    return gen_inputs


def test_config1_all_eight_4096_images(J, ref, gen):
    datas = [gen.c2_image(i) for i in range(8)]
    outs = J.decode_batch(datas, config=2)
    for i, (d, o) in enumerate(zip(datas, outs)):
        want = ref.decode_sampled(d, cfg=2)["pixels"].reshape(4096, 4096, 4)
        exact, mx = golden_lib.lossy_close(o.as_array(), want, "c2 image %d" % i, min_exact=0.999)
        assert mx <= 1


def test_config0_lossless_512_bit_exact(J, ref, gen):
    d = gen.c1_image()
    want = ref.decode_sampled(d, cfg=2)["pixels"]
    got = J.JxlCoder.decode(d, J.PreferredColorConfig.RGBA_8888)
    assert (got.width, got.height) == (512, 512)
    assert (got.pixels == want).all()
    # a batch of the same file decodes every copy identically
    for o in J.decode_batch([d] * 8, config=2):
        assert (o.pixels == want).all()


def test_config4_animation_frames(J, ref, gen):
    d = gen.c5_animation()
    a = J.JxlAnimatedImage(d, J.PreferredColorConfig.RGBA_8888)
    r = ref.Anim(d, cfg=2)
    assert a.number_of_frames == len(r) == 120
    assert (a.get_width(), a.get_height()) == r.size == (1024, 1024)
    frames = {}
    for i in range(120):   # every frame is decoded (in order, as a player would); every 9th is compared
        f = a.get_frame(i)
        assert (f.width, f.height) == (1024, 1024)
        assert a.get_frame_duration(i) == r.duration(i) == 40
        if i % 9 == 0 or i == 119:
            frames[i] = f.as_array().copy()
    for i, got in frames.items():
        want = r.frame(i)["pixels"].reshape(1024, 1024, 4)
        assert (got[..., 3] == want[..., 3]).all(), "alpha plane of frame %d" % i   # squeeze-coded alpha: integer path
        golden_lib.lossy_close(got, want, "c5 frame %d" % i, min_exact=0.995)
    a.close()
    r.close()


def test_config2_batch_of_256_1080p_f16(J, ref, gen):
    distinct = [gen.c3_image(i) for i in range(4)]
    outs = J.decode_batch([distinct[i % 4] for i in range(256)], config=3)
    wants = []
    for d in distinct:
        w = ref.decode_sampled(d, cfg=3)
        wants.append(np.ascontiguousarray(w["pixels"][:, : 1920 * 8]).view(np.float16).astype(np.float32))
    first = {}
    for i, o in enumerate(outs):
        assert (o.width, o.height, o.config) == (1920, 1080, "RGBA_F16")
        px = np.ascontiguousarray(o.pixels[:, : 1920 * 8])
        if i < 4:
            got = px.view(np.float16).astype(np.float32)
            d = np.abs(got - wants[i])
            # RGBA_F16 of an 8-bit source is half(u8 / 255): one 8-bit step is 1/255
            assert d.max() <= 1.0 / 255 + 1e-3, (i, float(d.max()))
            assert float((d == 0).mean()) >= 0.999, (i, float((d == 0).mean()))
            first[i] = px.copy()
        else:
            assert (px == first[i % 4]).all(), i   # every copy in the batch is the same picture


@pytest.mark.parametrize("prim,tf,name", [(0, 0, "SRGB"), (11, 13, "DISPLAY_P3"), (9, 16, "BT2020_PQ"), (1, 1, "BT2020_HLG"),
                                          (9, 1, "SRGB"), (11, 17, "DCI_P3")])
def test_color_space_tag_matches_reference(J, ref, prim, tf, name):
    """API >= 34: the Bitmap's ColorSpace.Named (JniDecoding.cpp:236-253, including its sRGB-primaries + 709 -> Hlg2100
    slip); below 34 no tag."""
    from oracle import synth
    img = synth.synth_image(96, 64, 5, alpha=False)
    data = ref.encode_ex(img, 96, 64, 3, bits=8, lossless=False, distance=1.0, alpha_distance=-1.0, options={"EFFORT": 3},
                         primaries=prim, transfer=tf, orientation=0)
    r = ref.decode_sampled(data, cfg=2, api_level=34)
    got = J.JxlCoder.decode(data, J.PreferredColorConfig.RGBA_8888)
    assert r["color_space"] == name, r["color_space"]
    assert got.color_space == r["color_space"]
    r33 = ref.decode_sampled(data, cfg=2, api_level=33)
    assert r33["color_space"] == ""


@pytest.mark.parametrize("seed,dist,effort", [(0, 1.0, 7), (1, 0.5, 3), (2, 2.0, 7), (3, 3.0, 5)])
def test_rec2100_pq_output(J, ref, seed, dist, effort):
    """PQ-tagged pictures are handed out as PQ (tag BT2020_PQ): libjxl's rational-polynomial PQ encode restated
    (numeric.h PqOetf); bound in golden_lib.pq_close."""
    from oracle import synth
    w, h = 640, 400
    img = synth.synth_image(w, h, seed, alpha=False)
    data = ref.encode_ex(img, w, h, 3, bits=8, lossless=False, distance=dist, alpha_distance=-1.0, options={"EFFORT": effort},
                         primaries=9, transfer=16, orientation=0)
    want = ref.decode_sampled(data, cfg=2)["pixels"].reshape(h, w, 4)
    got = J.JxlCoder.decode(data, J.PreferredColorConfig.RGBA_8888).as_array()
    golden_lib.pq_close(got, want, "PQ seed %d" % seed)
