"""GPU probe (not a test): configs[0] (512x512 lossless RGBA) latency, and a batch of 64 of them."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import jxl_coder_b200 as J  # noqa: E402
from oracle import gen_inputs  # noqa: E402

d = gen_inputs.c1_image()
J.JxlCoder.decode(d, 2)
ts = []
for _ in range(4):
    t = time.time()
    J.JxlCoder.decode(d, 2)
    ts.append(time.time() - t)
print("C1 single: best %.1f ms" % (min(ts) * 1e3), J.last_batch_timings())
ds = [d] * 64
J.decode_batch(ds, config=2)
t = time.time()
J.decode_batch(ds, config=2)
dt = time.time() - t
print("C1 batch of 64: %.1f ms -> %.1f MP/s" % (dt * 1e3, 64 * 0.262144 / dt))
