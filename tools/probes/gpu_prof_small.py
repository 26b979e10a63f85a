"""Tiny profiling workload for ncu: one prepared batch of N 4096x4096 images, run once."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import jxl_coder_b200 as J
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
datas = [open(os.path.join(ROOT, "bench_data", "c2_4096x4096_%02d.jxl" % (i % 8)), "rb").read() for i in range(n)]
b = J.PreparedBatch(datas, config=2)
for _ in range(runs):
    assert b.run() == 0
print(b.stage_ms())
b.free()
