"""GPU performance probe (not a test): stage times of a prepared batch of synthetic 4096x4096 lossy images."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import jxl_coder_b200 as J  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    files = sorted(f for f in os.listdir(os.path.join(ROOT, "bench_data")) if f.startswith("c2_%dx%d_" % (size, size)))
    datas = [open(os.path.join(ROOT, "bench_data", files[i % len(files)]), "rb").read() for i in range(n)]
    t = time.time()
    b = J.PreparedBatch(datas, config=2)
    print("prepare %.1f ms status %s" % ((time.time() - t) * 1e3, set(b.status)), flush=True)
    for r in range(reps):
        t = time.time()
        rc = b.run()
        dt = time.time() - t
        ms = b.stage_ms()
        mp = n * size * size / 1e6
        print("run %d rc=%d wall %.1f ms -> %.1f MP/s | %s" % (r, rc, dt * 1e3, mp / dt, {k: round(v, 2) for k, v in ms.items()}), flush=True)
    bmp = b.fetch(0)
    a = bmp.as_array()
    print("image0", a.shape, a[100, 100], a.mean())
    try:
        from oracle import refjxl
        if refjxl.available():
            want = refjxl.decode_sampled(datas[0], cfg=2)["pixels"].reshape(a.shape)
            d = np.abs(a.astype(int) - want.astype(int))
            print("vs reference: exact %.4f max %d" % ((d == 0).mean(), d.max()))
    except Exception as e:
        print("reference check skipped:", e)
    b.free()
    # one-shot e2e
    t = time.time()
    res = J.decode_batch(datas, config=2)
    dt = time.time() - t
    print("e2e decode_batch wall %.1f ms -> %.1f MP/s ; timings %s" % (dt * 1e3, n * size * size / 1e6 / dt, J.last_batch_timings()))
    t = time.time()
    res = J.decode_batch(datas, config=2)
    dt = time.time() - t
    print("e2e decode_batch (2nd) wall %.1f ms -> %.1f MP/s ; timings %s" % (dt * 1e3, n * size * size / 1e6 / dt, J.last_batch_timings()))


if __name__ == "__main__":
    main()
