"""Pre-generates every reference-encoded input the GPU tests use (tests/_cache, bench_data) so that the GPU box does not
spend its minutes encoding."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from oracle import gen_inputs, refjxl as ref, synth  # noqa: E402
import test_resize_host as T  # noqa: E402

for k in cases.SMALL:
    cases.get(k)
for kind in ("rgb_lossy", "rgba_lossless"):
    cases.anim_case(kind)
img = synth.synth_image(1024, 768, 11)
cases._cached("rgb_lossy_1024x768", lambda: ref.encode(img, 1024, 768))
for i in range(6):
    w, h = 512 + 64 * (i % 3), 384 + 40 * (i % 2)
    im = synth.synth_image(w, h, 20 + i)
    cases._cached("rgb_lossy_%dx%d_s%d" % (w, h, 20 + i), lambda: ref.encode(im, w, h))
for c in T.CASES:
    w, h = c[0], c[1]
    im = T._image(w, h, w * 1000 + h)
    cases._cached("resize_src_%dx%d" % (w, h), lambda: ref.encode(im[..., :3].reshape(-1), w, h, colorspace=1, compression=1))
print("c3", len(gen_inputs.c3_image(0)))
print("c4", len(gen_inputs.c4_image()))
