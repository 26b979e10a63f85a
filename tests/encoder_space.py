"""The option space of the reference's OWN encoder (JxlCoder.encode -> interop/JxlEncoding.cpp:140-165): quality (distance) or
lossless, effort 1..9 and decoding speed 0..4 -- the files a jxl-coder user is most likely to hand to decode()."""
import itertools

import cases
from oracle import synth

W, H = 300, 200
GRID = [(lossless, effort, ds) for lossless, effort, ds in itertools.product([True, False], [1, 4, 7, 9], [0, 2, 4])]


def name(lossless, effort, ds):
    return "encspace_%s_e%d_ds%d" % ("lossless" if lossless else "lossy", effort, ds)


def make(ref, lossless, effort, ds):
    img = synth.synth_image(W, H, 7, alpha=True)
    return cases._cached(name(lossless, effort, ds),
                         lambda: ref.encode_ex(img, W, H, 4, lossless=lossless, distance=1.0, options={"EFFORT": effort, "DECODING_SPEED": ds}))
