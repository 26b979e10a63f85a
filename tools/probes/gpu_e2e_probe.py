"""GPU probe (not a test): e2e submit / collect pipeline of 64 x 4096^2 batches with the library's timeline on stderr."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import jxl_coder_b200 as J  # noqa: E402


def main():
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 9
    files = sorted(f for f in os.listdir(os.path.join(ROOT, "bench_data")) if f.startswith("c2_4096x4096_"))
    datas = [open(os.path.join(ROOT, "bench_data", files[i % len(files)]), "rb").read() for i in range(64)]

    def run(n):
        inflight = []
        t0 = time.time()
        for i in range(n):
            ts = time.time()
            inflight.append(J.PendingBatch(datas, config=2, keep_native=True))
            sub = time.time() - ts
            if len(inflight) >= depth:
                tc = time.time()
                for b in inflight.pop(0).result():
                    b.free()
                print("step %d submit %.1f ms collect %.1f ms" % (i, sub * 1e3, (time.time() - tc) * 1e3), flush=True)
        for p in inflight:
            for b in p.result():
                b.free()
        return time.time() - t0
    run(2 * depth)
    dt = run(steps)
    print("depth %d: %.1f ms per step -> %.1f MP/s" % (depth, dt / steps * 1e3, 64 * 16.777216 * steps / dt), flush=True)


if __name__ == "__main__":
    main()
