"""GPU probe (not a test): the e2e submit / collect loop of bench.py run by K independent processes, one per GPU, on a
multi-GPU box -- isolates how the host side (cores, pinned memory, PCIe root) scales.  Usage: gpu_e2e_multi.py K [depth] [threads]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(dev, depth, steps):
    sys.path.insert(0, ROOT)
    import jxl_coder_b200 as J
    files = sorted(f for f in os.listdir(os.path.join(ROOT, "bench_data")) if f.startswith("c2_4096x4096_"))
    datas = [open(os.path.join(ROOT, "bench_data", files[i % len(files)]), "rb").read() for i in range(64)]

    def run(n):
        inflight = []
        t0 = time.time()
        for i in range(n):
            inflight.append(J.PendingBatch(datas, config=2, device=dev, keep_native=True))
            if len(inflight) >= depth:
                for b in inflight.pop(0).result():
                    b.free()
        for p in inflight:
            for b in p.result():
                b.free()
        return time.time() - t0
    run(2 * depth)
    # crude start barrier: wait for the next multiple of 5 s
    time.sleep(5 - time.time() % 5)
    dt = run(steps)
    print("gpu %d depth %d: %.1f ms per step -> %.2f GP/s" % (dev, depth, dt / steps * 1e3, 64 * 16.777216 * steps / dt / 1e3), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "child":
        child(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    else:
        k = int(sys.argv[1])
        depth = int(sys.argv[2]) if len(sys.argv) > 2 else 3
        threads = int(sys.argv[3]) if len(sys.argv) > 3 else max(2, (os.cpu_count() or 8) // k)
        env = dict(os.environ, JXLB_HOST_THREADS=str(threads), JXLB_PINNED_POOL_MB=str((depth + 2) * 4096 + 2048))
        ps = [subprocess.Popen([sys.executable, __file__, "child", str(g), str(depth), "12"], env=env) for g in range(k)]
        for p in ps:
            p.wait()
