"""CPU probe (not a test): randomised parity sweep of the decoder core through the host build of the kernel code
(tests/hostemu) against the reference -- sizes 1 .. 2400, RGB / RGBA, lossless and distances 0.1 .. 20, effort 1 .. 9, decoding
speed, EPF / Gaborish overrides, modular options, group order, progressive passes, resampling, colour encodings, four kinds of
picture (synthetic photo, noise, flat regions, gradient).  No GPU needed; what it cannot see is everything after the decoder
core (orientation, rescale, reformat), which tools/probes/gpu_sweep.py covers.

  python tools/probes/cpu_sweep.py SEED N          -> counts, every mismatch and refusal class
A fixed small instance runs in the CPU suite (tests/test_random_sweep_host.py)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def make_case(rng, k, seed):
    from oracle import synth
    w, h = int(rng.integers(1, 900)), int(rng.integers(1, 700))
    if rng.random() < 0.1:
        w, h = int(rng.integers(900, 2400)), int(rng.integers(200, 700))
    if rng.random() < 0.1:
        w, h = int(rng.integers(1, 40)), int(rng.integers(1, 40))
    alpha = rng.random() < 0.4
    ch = 4 if alpha else 3
    lossless = rng.random() < 0.35
    dist = float(rng.choice([0.1, 0.3, 0.5, 1.0, 1.5, 2.0, 3.0, 5.0, 8.0, 12.0, 20.0]))
    adist = float(rng.choice([-1.0, 0.0, 0.5, 1.0, 2.0, 4.0])) if alpha and not lossless else -1.0
    opts = {"EFFORT": int(rng.integers(1, 10))}
    if rng.random() < 0.4:
        opts["DECODING_SPEED"] = int(rng.integers(0, 5))
    if rng.random() < 0.15 and not lossless:
        opts["EPF"] = int(rng.integers(0, 4))
    if rng.random() < 0.15 and not lossless:
        opts["GABORISH"] = int(rng.integers(0, 2))
    if rng.random() < 0.1 and lossless:
        opts["MODULAR_GROUP_SIZE"] = int(rng.integers(0, 4))
    if rng.random() < 0.1 and lossless:
        opts["MODULAR_PREDICTOR"] = int(rng.integers(0, 16))
    if rng.random() < 0.1 and lossless:
        opts["MODULAR_NB_PREV_CHANNELS"] = int(rng.integers(0, 4))
    if rng.random() < 0.08:
        opts["GROUP_ORDER"] = 1
    if rng.random() < 0.12 and not lossless:
        opts[["PROGRESSIVE_AC", "QPROGRESSIVE_AC"][int(rng.integers(0, 2))]] = 1
    if rng.random() < 0.06 and not lossless:
        opts["RESAMPLING"] = 2
    prim, tf = [(0, 0), (0, 0), (0, 0), (11, 13), (9, 1), (1, 1)][int(rng.integers(0, 6))]
    kind = int(rng.integers(0, 4))
    if kind == 0:
        img = synth.synth_image(w, h, k + 1000 * seed, alpha=alpha)
    elif kind == 1:
        img = rng.integers(0, 256, (h, w, ch)).astype(np.uint8)
    elif kind == 2:
        img = (synth.synth_image(w, h, k, alpha=alpha).reshape(h, w, ch) // 64 * 64).astype(np.uint8)
    else:
        yy, xx = np.mgrid[0:h, 0:w]
        img = np.stack([xx * 255 // max(1, w - 1), yy * 255 // max(1, h - 1), (xx + yy) % 256] + ([(xx * yy) % 256] if alpha else []), axis=2).astype(np.uint8)
    desc = dict(k=k, w=w, h=h, ch=ch, lossless=lossless, dist=dist, adist=adist, opts=opts, prim=prim, tf=tf, kind=kind)
    return np.ascontiguousarray(img), desc


def run_case(ref, H, img, d):
    """-> ('ok' | 'refused' | 'bad', detail)"""
    w, h, ch = d["w"], d["h"], d["ch"]
    try:
        data = ref.encode_ex(img, w, h, ch, lossless=d["lossless"], distance=d["dist"], alpha_distance=d["adist"], options=d["opts"],
                             primaries=d["prim"], transfer=d["tf"])
        want = ref.decode_sampled(data, cfg=2)["pixels"][:, : w * 4].reshape(h, w, 4)
    except Exception:
        return "skip", "reference failed"
    try:
        e = H.Decoded(data)
    except RuntimeError as ex:
        return "refused", str(ex)
    if e.status:
        st = e.status
        e.close()
        return ("refused", "status %d" % st) if st in (3, 4) else ("bad", "status %d" % st)
    try:
        out = e.render()
    except RuntimeError as ex:
        e.close()
        return "refused", str(ex)
    late = e.late_status()
    e.close()
    if late:
        return "refused", "late status %d" % late
    if ch == 4:
        a = out[..., 3:4].astype(np.uint16)
        out[..., :3] = (out[..., :3].astype(np.uint16) * a // 255).astype(np.uint8)
    dd = np.abs(out.astype(int) - want.astype(int))
    if d["lossless"]:
        ok = dd.max() == 0
    else:
        # one 8-bit step; isolated samples (a dark channel of a saturated colour, premultiplied lossy alpha) may reach 2
        ok = dd.max() <= 2 and float((dd > 1).mean()) < 1e-5 and float((dd == 0).mean()) > 0.97
    if not ok:
        # the reference itself is not always deterministic on upsampled frames (tests/upsampling_cases.py: ref_decode_stable)
        for _ in range(4):
            again = ref.decode_sampled(data, cfg=2)["pixels"][:, : w * 4].reshape(h, w, 4)
            if not np.array_equal(again, want):
                return "skip", "reference not deterministic"
    return ("ok", "") if ok else ("bad", "max %d, exact %.4f, beyond one step %.2e" % (dd.max(), (dd == 0).mean(), (dd > 1).mean()))


def main():
    from oracle import refjxl as ref
    import hostemu_lib as H
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    rng = np.random.default_rng(seed)
    cnt = {"ok": 0, "refused": 0, "bad": 0, "skip": 0}
    why = {}
    t0 = time.time()
    for k in range(n):
        img, d = make_case(rng, k, seed)
        res, detail = run_case(ref, H, img, d)
        cnt[res] += 1
        if res == "refused":
            why[detail] = why.get(detail, 0) + 1
        if res == "bad":
            print("MISMATCH", detail, d, flush=True)
    print(seed, cnt, why, "%.0f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
