"""Frames coded at half resolution and upsampled 2x by the decoder (pixel_stages.h: StageUpsample2): what libjxl's encoder
emits at distances of about 10 and more (the reference's quality scale below ~25), or on request (RESAMPLING = 2)."""
import cases
from oracle import synth

GRID = [(w, h, dist, res, effort) for (w, h) in [(600, 400), (257, 255), (1101, 703), (97, 33), (16, 9)]
        for (dist, res, effort) in [(12.0, -1, 7), (20.0, -1, 7), (2.0, 2, 7), (25.0, -1, 3)]]


def name(w, h, dist, res, effort):
    return "up2_%dx%d_d%g_r%d_e%d" % (w, h, dist, res, effort)


def make(ref, w, h, dist, res, effort):
    img = synth.synth_image(w, h, 5)
    opts = {"EFFORT": effort}
    if res > 0:
        opts["RESAMPLING"] = res
    return cases._cached(name(w, h, dist, res, effort), lambda: ref.encode_ex(img, w, h, 3, distance=dist, options=opts))
