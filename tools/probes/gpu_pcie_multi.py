"""GPU probe (not a test): aggregate device->host bandwidth of plain pinned cudaMemcpyAsync on 1, 2, ... all visible GPUs at
once -- the ceiling of any host-output path on this box (bench.py's e2e moves 4 GiB of RGBA8 per 64-image step and GPU)."""
import threading
import time

import torch


def run(ngpu, seconds=2.0, mib=1024):
    bufs = []
    for g in range(ngpu):
        with torch.cuda.device(g):
            d = torch.empty(mib << 20, dtype=torch.uint8, device="cuda:%d" % g)
            h = [torch.empty(mib << 20, dtype=torch.uint8).pin_memory() for _ in range(2)]
            bufs.append((d, h, torch.cuda.Stream(device=g)))
    done = [0] * ngpu
    stop = [False]

    def work(g):
        d, h, s = bufs[g]
        with torch.cuda.device(g), torch.cuda.stream(s):
            k = 0
            while not stop[0]:
                h[k & 1].copy_(d, non_blocking=True)
                s.synchronize()
                done[g] += 1
                k += 1
    ths = [threading.Thread(target=work, args=(g,)) for g in range(ngpu)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    time.sleep(seconds)
    stop[0] = True
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    per = [c * mib / 1024 / dt for c in done]
    print("%d GPU(s) copying at once: aggregate %.1f GiB/s  (per GPU: %s)" % (ngpu, sum(per), ", ".join("%.1f" % p for p in per)), flush=True)


if __name__ == "__main__":
    n = torch.cuda.device_count()
    k = 1
    while k <= n:
        run(k)
        k *= 2
