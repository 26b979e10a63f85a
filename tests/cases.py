"""Seeded synthetic inputs shared by the CPU and GPU parity tests.  Encoded with the reference's own encoder through
oracle/_ref and cached under tests/_cache (git-ignored; the cache travels to the GPU box with the snapshot)."""
import hashlib
import os
import numpy as np

from oracle import refjxl, synth

CACHE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_cache")


def natural_like(w, h, seed=5):
    """Brownian-noise field with flat patches and hard edges: drives the encoder into the special 8x8 transforms
    (IDENTITY, DCT2X2, DCT4X4, AFV, DCT4X8) that smooth synthetic images never use."""
    rng = np.random.default_rng(seed)
    a = np.cumsum(np.cumsum(rng.standard_normal((h, w, 3)), 0), 1) * 2 + 128 + rng.standard_normal((h, w, 3)) * 3
    a = np.clip(a, 0, 255).astype(np.uint8)
    a[h // 4:h // 4 + h // 3, w // 6:w // 6 + w // 3] = (a[h // 4:h // 4 + h // 3, w // 6:w // 6 + w // 3] // 32) * 32
    a[h // 2:h // 2 + 1, :, :] = 255
    a[:, (w * 3) // 8:(w * 3) // 8 + 2, :] = 0
    return a


def _cached(name, make):
    os.makedirs(CACHE, exist_ok=True)
    p = os.path.join(CACHE, name + ".jxl")
    if os.path.exists(p):
        return open(p, "rb").read()
    data = make()
    with open(p, "wb") as f:
        f.write(data)
    return data


# name -> (callable producing the encoded bytes)
def _cases():
    c = {}
    c["rgb_lossy_64"] = lambda: refjxl.encode(synth.synth_image(64, 64, 4), 64, 64)
    c["rgb_lossy_64_epf3"] = lambda: refjxl.encode_ex(synth.synth_image(64, 64, 4), 64, 64, 3, options={"EPF": 3})
    c["rgb_lossy_64_epf2_nogab"] = lambda: refjxl.encode_ex(synth.synth_image(64, 64, 4), 64, 64, 3, options={"EPF": 2, "GABORISH": 0})
    c["rgb_lossy_64_epf0"] = lambda: refjxl.encode_ex(synth.synth_image(64, 64, 4), 64, 64, 3, options={"EPF": 0})
    c["rgb_lossy_64_d03"] = lambda: refjxl.encode_ex(synth.synth_image(64, 64, 4), 64, 64, 3, distance=0.3)
    c["rgb_lossy_64_d3"] = lambda: refjxl.encode_ex(synth.synth_image(64, 64, 4), 64, 64, 3, distance=3.0)
    c["rgb_lossy_256x200"] = lambda: refjxl.encode(synth.synth_image(256, 200, 0), 256, 200)
    c["rgb_lossy_256x200_e3"] = lambda: refjxl.encode(synth.synth_image(256, 200, 0), 256, 200, effort=3)
    c["rgba_lossy_300x203"] = lambda: refjxl.encode(synth.synth_image(300, 203, 1, alpha=True), 300, 203, colorspace=2)
    c["rgb_lossy_2304x24"] = lambda: refjxl.encode(synth.synth_image(2304, 24, 2), 2304, 24)
    c["rgb_lossy_24x2100"] = lambda: refjxl.encode(synth.synth_image(24, 2100, 2), 24, 2100)
    c["rgba_lossless_128"] = lambda: refjxl.encode(synth.synth_image(128, 128, 3, alpha=True), 128, 128, colorspace=2, compression=1)
    c["rgba_lossless_300x260"] = lambda: refjxl.encode(synth.synth_image(300, 260, 3, alpha=True), 300, 260, colorspace=2, compression=1)
    c["rgb_lossless_200x150"] = lambda: refjxl.encode(synth.synth_image(200, 150, 6), 200, 150, colorspace=1, compression=1)
    for dist in (0.5, 1.0, 2.0, 4.0):
        for eff in (5, 7):
            c["natural_d%g_e%d" % (dist, eff)] = (lambda d=dist, e=eff: refjxl.encode_ex(natural_like(200, 200), 200, 200, 3, distance=d, options={"EFFORT": e}))
    c["natural_512_d1"] = lambda: refjxl.encode_ex(natural_like(512, 384, 9), 512, 384, 3, distance=1.0)
    return c


CASES = _cases()
SMALL = [k for k in CASES]


def get(name):
    return _cached(name, CASES[name])


# ---- animations (the reference's JxlAnimatedEncoder: full-canvas kReplace frames, 40 ms each) ----
ANIM_W, ANIM_H, ANIM_N = 320, 264, 4


def anim_case(kind):
    """kind: "rgb_lossy" | "rgba_lossless" | "rgba_lossy" (alpha coded with the squeeze transform)."""
    alpha = kind in ("rgba_lossless", "rgba_lossy")

    def make():
        ch = 4 if alpha else 3
        frames = np.stack([synth.synth_image(ANIM_W, ANIM_H, 40 + i, alpha=alpha).reshape(ANIM_H, ANIM_W, ch) for i in range(ANIM_N)])
        return refjxl.anim_encode(frames, ANIM_W, ANIM_H, colorspace=2 if alpha else 1, compression=1 if kind == "rgba_lossless" else 2, duration=40)
    return _cached("anim_%s_%dx%dx%d" % (kind, ANIM_W, ANIM_H, ANIM_N), make)
