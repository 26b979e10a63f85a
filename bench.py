#!/usr/bin/env python
"""Benchmark of the JXL -> RGBA8 decode path (BASELINE.json: "Mpixels/s decoded (JXL->RGBA8)"; workload = configs[1]:
batch of 64 synthetic 4096x4096 lossy VarDCT (q=90) images -> RGBA_8888 per GPU).

    python bench.py --gpus N --steps K --warmup W                 # our CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W # the reference's own CPU path (oracle/_ref), rank 0 only

A step = one pass of the whole hot path over one batch.  `value` times the kernels with the codestreams + parsed tables
already resident in HBM (jxlb_batch_prepare once, jxlb_batch_run per step); `e2e` times jxlb_decode_batch on host
buffers, i.e. header parsing, pinned staging, H2D, all kernels and the D2H of every decoded pixel.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIZE = 4096
BATCH = 64
DISTINCT = 8
MPIX_PER_IMAGE = SIZE * SIZE / 1e6
# algorithmic bytes per pixel of the inverse-transform kernel (SURVEY.md §8d "if split: K_idct"):
# 3 x int16 coefficients read (6 B) + per-cell metadata and LF (0.25 B) + 3 x f32 XYB samples written (12 B)
IDCT_BYTES_PER_PIXEL = 18.25
# dram__bytes_read.sum + dram__bytes_write.sum of the inverse-transform launches of one 4096x4096 image (ReconRegionTmaKernel
# 105.4 + 143.1 MB, ReconLargeListKernel 0.07 MB), ncu --set full of this HEAD's kernel: profiles/r2_ncu_recon.txt
TRAFFIC_BYTES_PER_IMAGE = 248_600_000


def load_inputs(batch=BATCH, distinct=DISTINCT, size=SIZE):
    """`distinct` different synthetic images (seeds 12345+i, oracle/synth.py) encoded with the reference's encoder
    settings and cached under bench_data/; the batch cycles through them."""
    from oracle import gen_inputs
    datas = [gen_inputs.c2_image(i, size) for i in range(distinct)]
    return [datas[i % distinct] for i in range(batch)]


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of this rank's GPU while the timed region runs: NVML in-process (a query costs
    microseconds); only when the nvidia_ml_py module is missing does it fall back to spawning nvidia-smi, once a second --
    one nvidia-smi process per rank every 150 ms kept several host cores and the driver's locks busy on 4 / 8 GPU boxes."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def _nvml_loop(self):
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(self.index)
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]
        while not self.stop_flag:
            try:
                r = get_reasons(h)
                self.samples.append([str(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), str(mx)] +
                                    ["Active" if r & b else "Not Active" for _, b in bits])
            except Exception:
                pass
            time.sleep(0.2)

    def run(self):
        try:
            self._nvml_loop()
            return
        except Exception:
            pass
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.samples.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(1.0)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for k, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(k)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.samples)}


def cpu_reference(datas, threads, seconds_budget, per_thread_images):
    """Times the reference's DecodeJpegXlOneShot (oracle/_ref: its own sources + prebuilt libjxl 0.12.0) on host cores:
    `threads` concurrent decodes (each creating libjxl's resizable thread pool exactly as the reference does)."""
    from oracle import refjxl
    refjxl.lib()
    refjxl.decode_discard(datas[0])  # warm up
    done = [0] * threads

    def work(t):
        for i in range(per_thread_images):
            refjxl.decode_discard(datas[(t + i) % len(datas)])
            done[t] += 1
    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.time()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.time() - t0
    return sum(done) * MPIX_PER_IMAGE / dt, dt, sum(done)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refjxl
    if not refjxl.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref was not built (needs /root/reference once)"}))
        return
    datas = load_inputs()
    cores = os.cpu_count() or 1
    # one step = a bounded sample of the 64-image workload: `cores` concurrent decodes x 1 image each
    sample_images = 2 * max(cores, 8)
    vals = []
    for i in range(args.warmup):
        cpu_reference(datas, cores, 0, 1)
    t0 = time.time()
    for i in range(args.steps):
        v, dt, nimg = cpu_reference(datas, cores, 0, max(1, sample_images // cores))
        vals.append(v)
    total = time.time() - t0
    value = sum(vals) / len(vals)
    out = {
        "impl": "reference", "metric": "Mpixels/s decoded (JXL->RGBA8)", "value": round(value, 2), "unit": "MP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * total / args.steps, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "batch of 64 synthetic 4096x4096 lossy VarDCT (q=90) JXL -> RGBA8 (configs[1]); reference arm decodes a bounded sample per step",
                   "sample_images_per_step": sample_images},
        "cpu_baseline": {"value": round(value, 2), "unit": "MP/s", "cores": cores, "kind": "reference",
                         "sample": "%d concurrent DecodeJpegXlOneShot calls (libjxl 0.12.0 prebuilt x86_64, SSE2 build, own thread pool per call) on %d 4096x4096 images per step" % (cores, sample_images)},
        "e2e": {"value": round(value, 2), "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the decode path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # the library's pinned-buffer pool and host thread count are per process: share the host among the ranks.  The pool
    # must hold every result buffer in flight (depth batches of 4 GiB) plus the batch the caller is still reading and one
    # being recycled: a smaller cap turns every step into a 4 GiB cudaHostAlloc / cudaFreeHost pair (~1 s of page pinning)
    depth = e2e_depth(args, world)
    os.environ.setdefault("JXLB_PINNED_POOL_MB", str((depth + 2) * 4096 + 2048))
    os.environ.setdefault("JXLB_HOST_THREADS", str(max(2, (os.cpu_count() or 8) // world)))
    import jxl_coder_b200 as J
    J.load_library()
    datas = load_inputs()
    h2d = sum(len(d) for d in datas)
    d2h = BATCH * SIZE * SIZE * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident measurement: CONTEXTS prepared batches (each a full 64-image batch with its own buffers and
    # stream) are run alternately, one batch run = one step, so the latency-bound LF-group stage of step k+1 overlaps the
    # throughput kernels of step k.  Timed on the device: CUDA events from the first run's start to the last run's end.
    contexts = max(1, min(args.contexts, args.steps))
    batches = [J.PreparedBatch(datas, config=J.PreferredColorConfig.RGBA_8888, device=local) for _ in range(contexts)]
    for b in batches:
        assert all(s == 0 for s in b.status), b.status
    launches0 = J.kernel_launches()
    for i in range(args.warmup):
        assert batches[i % contexts].run_async() == 0
    for b in batches[:min(contexts, args.warmup)]:
        assert b.wait() == 0
    launches_per_step = (J.kernel_launches() - launches0) // max(1, args.warmup)
    for b in batches:
        b.reset_stats()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        assert batches[i % contexts].run_async() == 0
    for b in batches:
        assert b.wait() == 0
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = max(batches[0].span_ms(b) for b in batches)
    stage_acc, stage_n = {}, 0
    for b in batches:
        ms, nruns = b.stage_ms_mean()
        for k, v in ms.items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v * nruns
        stage_n += nruns
    assert stage_n == args.steps, (stage_n, args.steps)
    batch = batches[(args.steps - 1) % contexts]
    # roofline pass: the same step run ALONE (one context, nothing overlapping), so that the CUDA-event time of the
    # dominant kernel is its own duration rather than its share of a GPU it splits with three other batches
    alone_ms = {}
    if args.steps > 0:
        batch.reset_stats()
        for _ in range(2):
            assert batch.run() == 0
        alone_ms, _ = batch.stage_ms_mean()
    # correctness spot check of what was just timed (against the reference when it is present on this box)
    parity = None
    try:
        from oracle import refjxl
        if rank == 0 and refjxl.available():
            import numpy as np
            got = batch.fetch(0).as_array()
            want = refjxl.decode_sampled(datas[0], cfg=2)["pixels"].reshape(got.shape)
            d = np.abs(got.astype(np.int32) - want.astype(np.int32))
            parity = {"exact": round(float((d == 0).mean()), 4), "max_abs_diff": int(d.max())}
    except Exception as e:  # the check is informative only
        parity = {"error": str(e)}
    for b in batches:
        b.free()
    # device time (CUDA events on the decode streams), max over ranks
    t_dev = torch.tensor([dev_ms / 1e3], device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_s = float(t_dev.item())
    value = world * BATCH * MPIX_PER_IMAGE * args.steps / dev_s

    # ---- end to end through the public batch API: host buffers in, pinned host pixels out, every step includes header
    # parsing, pinned staging, H2D of the codestreams, all kernels and the D2H of every decoded pixel (4 GiB per step).
    # ONE caller thread keeps `depth` batches in flight with jxlb_decode_batch_submit / _collect (a synchronous call is a
    # latency chain -- its LF stage alone is ~86 ms of serial entropy chains -- so a single call at a time leaves the GPU
    # and the PCIe link idle most of the time; the reference's own callers overlap calls from worker pools).
    # every batch in flight keeps 4 GiB of pinned result buffers: on 4 / 8 GPUs (ranks share the host's RAM and cores) fewer

    def e2e_steps_run(nsteps):
        inflight = []
        for i in range(nsteps):
            inflight.append(J.PendingBatch(datas, config=2, device=local, keep_native=True))
            if len(inflight) >= depth:
                for b in inflight.pop(0).result():
                    b.free()
        for p in inflight:
            for b in p.result():
                b.free()

    def e2e_sync_run(nsteps, callers):
        nxt = [0]
        lock = threading.Lock()
        errs = []

        def work():
            while True:
                with lock:
                    if nxt[0] >= nsteps:
                        return
                    nxt[0] += 1
                try:
                    for b in J.decode_batch(datas, config=2, device=local, keep_native=True):
                        b.free()
                except Exception as e:  # noqa
                    errs.append(e)
                    return
        ths = [threading.Thread(target=work) for _ in range(min(callers, nsteps))]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        if errs:
            raise errs[0]

    e2e_steps_run(2 * depth if args.warmup else 0)  # every decode slot allocates its buffers + pinned pool once
    e2e_steps = max(1, args.steps, 6 * depth)

    def collect(p):
        for b in p.result():
            b.free()

    # (1) cold: K steps from an empty pipeline until the last batch is on the host (includes filling and draining it)
    barrier()
    t1 = time.perf_counter()
    e2e_steps_run(e2e_steps)
    barrier()
    t_cold = torch.tensor([time.perf_counter() - t1], device="cuda")
    # (2) steady state (the headline): the pipeline is primed with `depth` batches before the timer starts and is still
    # full when it stops; exactly K batches are collected inside the timed region, each followed by a new submission
    inflight = [J.PendingBatch(datas, config=2, device=local, keep_native=True) for _ in range(depth)]
    collect(inflight.pop(0))
    inflight.append(J.PendingBatch(datas, config=2, device=local, keep_native=True))
    barrier()
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        collect(inflight.pop(0))
        inflight.append(J.PendingBatch(datas, config=2, device=local, keep_native=True))
    e2e_wall = time.perf_counter() - t1
    for p_ in inflight:
        collect(p_)
    barrier()
    t_e2e = torch.tensor([e2e_wall], device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_cold, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * MPIX_PER_IMAGE * e2e_steps / float(t_e2e.item())
    e2e_cold_value = world * BATCH * MPIX_PER_IMAGE * e2e_steps / float(t_cold.item())
    # secondary: plain synchronous jxlb_decode_batch calls from 2 caller threads
    sync_callers = 2
    e2e_sync_run(2, sync_callers)
    barrier()
    t2 = time.perf_counter()
    sync_steps = 8
    e2e_sync_run(sync_steps, sync_callers)
    barrier()
    t_sync = torch.tensor([time.perf_counter() - t2], device="cuda")
    if world > 1:
        dist.all_reduce(t_sync, op=dist.ReduceOp.MAX)
    e2e_sync_value = world * BATCH * MPIX_PER_IMAGE * sync_steps / float(t_sync.item())
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if rank != 0:
        return
    steps = args.steps
    stages = {k: round(v / max(1, stage_n), 3) for k, v in stage_acc.items()}
    idct_ms = alone_ms.get("inverse_transforms", 0.0) or stages["inverse_transforms"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = IDCT_BYTES_PER_PIXEL * BATCH * SIZE * SIZE / (idct_ms * 1e-3) / 1e9 if idct_ms > 0 else 0.0
    out = {
        "metric": "Mpixels/s decoded (JXL->RGBA8)", "value": round(value, 1), "unit": "MP/s", "n_gpus": world, "steps": steps,
        "warmup": args.warmup, "ms_per_step": round(1e3 * dev_s / steps, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "batch of 64 synthetic 4096x4096 lossy VarDCT (q=90, effort 7) JXL -> RGBA_8888 per GPU (configs[1]); %d distinct images cycled" % DISTINCT,
                   "images_per_gpu": BATCH, "l2": "inputs_larger_than_L2 (each step touches > 30 GB of planes)",
                   "value_is": "kernels only, codestreams + host-parsed tables resident in HBM; %d prepared batches (decode contexts) run alternately with jxlb_batch_run_async, one batch run per step; device time = CUDA events first-run start -> last-run end" % contexts,
                   "contexts": contexts,
                   "stages_note": "stages_ms_per_step are per-run CUDA-event intervals on the run's own stream; with 2 contexts they include time shared with the other context's kernels",
                   "roofline_kernel": "ReconRegionTmaKernel + ReconLargeListKernel (dequant + CfL + LLF + inverse VarDCT -> XYB f32 planes)",
                   "wall_ms_per_step": round(1e3 * wall / steps, 2)},
        "e2e": {"value": round(e2e_value, 1), "unit": "MP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "ms_per_step": round(1e3 * float(t_e2e.item()) / e2e_steps, 2),
                "callers": 1, "in_flight": depth,
                "timing": "steady state: the pipeline is primed with in_flight batches before the timer starts and still full when it stops; K = steps batches are collected inside the timed region (wall clock, max over ranks), each collect followed by the next submit; every batch's parse, H2D, kernels and D2H run inside the pipeline",
                "cold": {"value": round(e2e_cold_value, 1), "unit": "MP/s", "ms_per_step": round(1e3 * float(t_cold.item()) / e2e_steps, 2),
                         "timing": "the same K steps from an empty pipeline until the last batch is on the host (adds one batch latency of filling and draining)"},
                "pcie_ceiling": "4 GiB of RGBA8 per step over a PCIe Gen5 x16 link measured at 57 GB/s = 75 ms per step = 14.3 GP/s per GPU",
                "api": "jxlb_decode_batch_submit / _collect (host buffers -> pinned host RGBA), 1 caller thread, %d batches in flight" % depth,
                "sync_2_callers": {"value": round(e2e_sync_value, 1), "unit": "MP/s", "steps": sync_steps,
                                   "api": "synchronous jxlb_decode_batch from 2 caller threads"}},
        "gpu_launches": int(launches_per_step * steps),
        "stages_ms_per_step": stages,
        "stages_ms_alone": {k: round(v, 3) for k, v in alone_ms.items()},
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": TRAFFIC_BYTES_PER_IMAGE * BATCH if TRAFFIC_BYTES_PER_IMAGE else None,
                     "kernel_ms": round(idct_ms, 3), "launches": 2 * BATCH,
                     "how": "64 ReconRegionTmaKernel + 64 ReconLargeListKernel launches of one 64-image batch run alone (no other context in flight) after the timed region; every 4th image's launch pair is bracketed by CUDA events on its stream, mean x 64; peak = burst HBM copy bandwidth",
                     "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                     "algorithmic_bytes_per_pixel": IDCT_BYTES_PER_PIXEL},
        "clocks": sampler.summary(),
        "parity_spot_check": parity,
    }
    # bounded CPU baseline on this box (rank 0, N=1 only)
    if world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import refjxl
            if refjxl.available():
                cores = os.cpu_count() or 1
                v, dt, nimg = cpu_reference(datas, cores, 0, 2)
                out["cpu_baseline"] = {"value": round(v, 2), "unit": "MP/s", "cores": cores, "kind": "reference",
                                       "sample": "%d concurrent DecodeJpegXlOneShot calls (reference's prebuilt libjxl 0.12.0, SSE2) on %d of the 4096x4096 images, %.1f s" % (cores, nimg, dt)}
            else:
                out["cpu_baseline"] = {"value": None, "unit": "MP/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not present on this box"}
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "unit": "MP/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % e}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the decode path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ.setdefault("JXLB_PINNED_POOL_MB", str(max(8192, 40960 // world)))
    os.environ.setdefault("JXLB_HOST_THREADS", str(max(2, (os.cpu_count() or 8) // world)))
    return torch, dist, world, rank, local


def _timed_steps(torch, dist, world, step, steps, warmup):
    """W warm-up steps, then K timed steps between barrier + synchronize; returns seconds (max over ranks)."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(warmup):
        step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    barrier()
    t = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_c3(args):
    """configs[2]: 256 x 1080p lossy -> RGBA_F16, the batch sharded over the ranks by compressed size (strong scaling: the
    256 images are the whole job at every N).  Host codestreams in, pinned host RGBA_F16 out (e2e through jxlb_decode_batch)."""
    torch, dist, world, rank, local = _dist_setup()
    import jxl_coder_b200 as J
    from jxl_coder_b200 import shard
    from oracle import gen_inputs
    J.load_library()
    distinct = [gen_inputs.c3_image(i) for i in range(4)]
    datas = [distinct[i % 4] for i in range(256)]
    idx, mine = shard.shard_for_rank(datas, world, rank)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = J.kernel_launches()

    def step():
        for b in J.decode_batch(mine, config=3, device=local, keep_native=True):
            b.free()
    secs = _timed_steps(torch, dist, world, step, args.steps, max(3, args.warmup))
    launches = J.kernel_launches() - launches0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if rank != 0:
        return
    mp = 256 * 1920 * 1080 / 1e6
    value = mp * args.steps / secs
    out = {"metric": "Mpixels/s decoded (JXL->RGBA_F16)", "value": round(value, 1), "unit": "MP/s", "n_gpus": world, "steps": args.steps,
           "warmup": max(3, args.warmup), "ms_per_step": round(1e3 * secs / args.steps, 2), "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "batch of 256 synthetic 1920x1080 lossy VarDCT (q=90) JXL -> RGBA_F16, sharded over the GPUs by compressed size (configs[2]); 4 distinct images cycled",
                      "images_per_gpu": len(mine), "l2": "inputs_larger_than_L2 (each step touches > 9 GB of planes)"},
           "e2e": {"value": round(value, 1), "unit": "MP/s", "h2d_bytes_per_step": sum(len(d) for d in datas), "d2h_bytes_per_step": 256 * 1920 * 1080 * 8,
                   "api": "jxlb_decode_batch (host buffers -> pinned host RGBA_F16), one synchronous call per rank per step"},
           "gpu_launches": int(launches), "clocks": sampler.summary()}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_c5(args):
    """configs[4]: the 120-frame 1024x1024 RGBA lossy animation, every frame through JxlAnimatedImage.getFrame; the frames
    are dealt to the ranks in blocks of 32 (the handle's prefetch unit): each rank opens the file and asks for its frames in order."""
    torch, dist, world, rank, local = _dist_setup()
    import jxl_coder_b200 as J
    from oracle import gen_inputs
    J.load_library()
    data = gen_inputs.c5_animation()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = J.kernel_launches()

    def step():
        a = J.JxlAnimatedImage(data, J.PreferredColorConfig.RGBA_8888)
        n = a.number_of_frames
        block = 32  # the handle decodes 32 frames per batch and reads one block ahead: blocks are dealt round-robin to the ranks
        for first in range(rank * block, n, world * block):
            for i in range(first, min(first + block, n)):
                a.get_frame(i)
        a.close()
    secs = _timed_steps(torch, dist, world, step, args.steps, max(3, args.warmup))
    launches = J.kernel_launches() - launches0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if rank != 0:
        return
    mp = 120 * 1024 * 1024 / 1e6
    value = mp * args.steps / secs
    out = {"metric": "Mpixels/s decoded (JXL->RGBA8)", "value": round(value, 1), "unit": "MP/s", "n_gpus": world, "steps": args.steps,
           "warmup": max(3, args.warmup), "ms_per_step": round(1e3 * secs / args.steps, 2), "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "JxlAnimatedImage: 120-frame 1024x1024 RGBA lossy animation, getFrame(i) for every frame, blocks of 32 frames dealt round-robin to the GPUs (configs[4])",
                      "ms_per_frame": round(1e3 * secs / args.steps / 120, 3)},
           "e2e": {"value": round(value, 1), "unit": "MP/s", "h2d_bytes_per_step": len(data), "d2h_bytes_per_step": 120 * 1024 * 1024 * 4,
                   "api": "jxlb_anim_open + jxlb_anim_get_frame per frame (host file -> pinned host RGBA8)"},
           "gpu_launches": int(launches), "clocks": sampler.summary()}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def e2e_depth(args, world):
    """Batches one caller keeps in flight: 5 on one GPU, fewer when several ranks share the host's RAM and cores."""
    return max(1, args.depth if world == 1 else min(args.depth, 4) if world <= 4 else min(args.depth, 3))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--contexts", type=int, default=int(os.environ.get("JXLB_BENCH_CONTEXTS", "6")),
                    help="prepared batches (decode contexts) alternated by the device-resident measurement")
    ap.add_argument("--depth", type=int, default=int(os.environ.get("JXLB_BENCH_DEPTH", "5")),
                    help="batches the e2e measurement keeps in flight (jxlb_decode_batch_submit / _collect, one caller thread)")
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c5"],
                    help="c2 = BASELINE configs[1] (the headline, default); c3 = configs[2] (256 x 1080p -> RGBA_F16 sharded); c5 = configs[4] (animation)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "c3":
        run_c3(args)
    elif args.config == "c5":
        run_c5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
