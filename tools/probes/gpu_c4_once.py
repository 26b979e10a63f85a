"""GPU probe (not a test): configs[3] once -- 8K -> 1080p RGBA_1010102 at API level 33 (orientation-free; rescale, colour
pass and reformat kernels all run).  Used under ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import jxl_coder_b200 as J  # noqa: E402
from oracle import gen_inputs  # noqa: E402

d = gen_inputs.c4_image()
J.JxlCoder.api_level = 33
for _ in range(2):
    b = J.JxlCoder.decode_sampled(d, 1920, 1080, 5, 1, 4)
print(b.width, b.height, b.config, J.last_batch_timings())
