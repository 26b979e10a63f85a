"""TEST/BENCH INFRASTRUCTURE — generates (and caches under bench_data/, git-ignored) the synthetic inputs of the
BASELINE.json configs with the reference's own encoder settings (oracle/synth.py + oracle/refjxl.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import refjxl, synth  # noqa: E402

DATA = os.path.join(ROOT, "bench_data")


def cached(name, make):
    os.makedirs(DATA, exist_ok=True)
    p = os.path.join(DATA, name)
    if os.path.exists(p):
        return open(p, "rb").read()
    d = make()
    with open(p + ".tmp", "wb") as f:
        f.write(d)
    os.replace(p + ".tmp", p)
    return d


def c2_image(index, size=4096):
    """configs[1]: 4096x4096 RGB lossy q=90 (distance 1.0), effort 7."""
    return cached("c2_%dx%d_%02d.jxl" % (size, size, index), lambda: refjxl.encode(synth.synth_image(size, size, index), size, size))


def c1_image():
    """configs[0]: 512x512 lossless RGBA8."""
    return cached("c1_512_lossless_rgba.jxl", lambda: refjxl.encode(synth.synth_image(512, 512, 0, alpha=True), 512, 512, colorspace=2, compression=1))


def c3_image(index):
    """configs[2]: 1920x1080 RGB lossy."""
    return cached("c3_1080p_%02d.jxl" % index, lambda: refjxl.encode(synth.synth_image(1920, 1080, 100 + index), 1920, 1080))


def c4_image():
    """configs[3]: 7680x4320 RGB lossy."""
    return cached("c4_8k.jxl", lambda: refjxl.encode(synth.synth_image(7680, 4320, 200), 7680, 4320))


def c5_animation(frames=120, size=1024):
    """configs[4]: 120-frame 1024x1024 RGBA lossy animation written like the reference's JxlAnimatedEncoder (40 ms frames)."""
    def make():
        fr = np.stack([synth.synth_image(size, size, 300 + i, alpha=True).reshape(size, size, 4) for i in range(frames)])
        return refjxl.anim_encode(fr, size, size, colorspace=2, compression=2, duration=40)
    return cached("c5_anim_%dx%dx%d.jxl" % (size, size, frames), make)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    for i in range(n):
        d = c2_image(i)
        print("c2", i, len(d), flush=True)
    print("c1", len(c1_image()))
    print("c2 small", len(c2_image(0, 2048)), len(c2_image(0, 1024)))
