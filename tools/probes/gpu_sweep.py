"""GPU probe (not a test): randomised parity sweep against the reference over encoder settings -- sizes, distances,
efforts, alpha (lossless / lossy), orientation, colour encodings, bit depth, lossless, decodeSampled arguments.
Every case ends as MATCH, REFUSED (UnsupportedJXLException) or MISMATCH; the last kind is printed in full."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import jxl_coder_b200 as J  # noqa: E402
from oracle import refjxl as ref, synth  # noqa: E402


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 120
    rng = np.random.default_rng(seed)
    # MATCH_PQ_DARK: Rec.2100-PQ output whose only samples beyond the tolerance sit where the reference is in the darkest
    # quarter of the range and are rarer than 3 in 10^5 (1 in 10^3 after a rescale has spread them): the bound of
    # tests/golden_lib.pq_close, the PQ curve being ill-conditioned at black.
    counts = {"MATCH": 0, "MATCH_PQ_DARK": 0, "REFUSED": 0, "MISMATCH": 0, "REF_ERR": 0}
    refused = {}
    only = os.environ.get("SWEEP_ONLY_LOSSLESS")
    for k in range(n):
        w, h = int(rng.integers(1, 700)), int(rng.integers(1, 600))
        if rng.random() < 0.15:
            w, h = int(rng.integers(600, 1400)), int(rng.integers(300, 900))
        alpha = rng.random() < 0.4
        ch = 4 if alpha else 3
        lossless = rng.random() < 0.3
        bits = 8 if rng.random() < 0.8 else 16
        dist = float(rng.choice([0.3, 0.5, 1.0, 1.5, 2.0, 3.0, 5.0, 8.0]))
        adist = float(rng.choice([-1.0, 0.0, 0.5, 1.0, 2.0])) if alpha and not lossless else -1.0
        effort = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 8]))
        orient = int(rng.choice([0, 0, 0, 2, 3, 4, 5, 6, 7, 8]))
        prim, tf = [(0, 0), (0, 0), (11, 13), (9, 1), (1, 1), (9, 16), (1, 8)][int(rng.integers(0, 7))]
        opts = {"EFFORT": effort}
        if rng.random() < 0.2:
            opts["EPF"] = int(rng.integers(0, 4))
        if rng.random() < 0.2:
            opts["GABORISH"] = int(rng.integers(0, 2))
        if bits == 16:
            img = (synth.synth_image(w, h, k, alpha=alpha).astype(np.uint16) * 257)
        else:
            img = synth.synth_image(w, h, k, alpha=alpha)
        desc = dict(w=w, h=h, ch=ch, lossless=lossless, bits=bits, dist=dist, adist=adist, opts=opts, orient=orient, prim=prim, tf=tf)
        cfg = int(rng.choice([2, 2, 3, 4, 5, 1]))
        sampled = rng.random() < 0.3
        args = (-1, -1, cfg, 1, 4)
        if sampled:
            args = (int(rng.integers(1, 400)), int(rng.choice([-1, -2, int(rng.integers(1, 400))])), cfg, int(rng.integers(1, 4)),
                    int(rng.choice([1, 2, 3, 4, 6, 7, 8])))
        desc["args"] = args
        if only and not lossless:
            continue
        try:
            data = ref.encode_ex(img, w, h, ch, bits=bits, lossless=lossless, distance=dist, alpha_distance=adist, options=opts,
                                 primaries=prim, transfer=tf, orientation=orient)
        except Exception as e:
            counts["REF_ERR"] += 1
            continue
        try:
            r = ref.decode_sampled(data, w=args[0], h=args[1], cfg=args[2], scale_mode=args[3], filt=args[4])
        except Exception as e:
            counts["REF_ERR"] += 1
            continue
        try:
            got = J.JxlCoder.decode_sampled(data, *args)
        except J.UnsupportedJXLException as e:
            counts["REFUSED"] += 1
            refused[str(e)] = refused.get(str(e), 0) + 1
            if os.environ.get("SWEEP_SHOW_REFUSED") and os.environ["SWEEP_SHOW_REFUSED"] in str(e):
                print("REFUSED", str(e), desc, flush=True)
            continue
        except Exception as e:
            counts["MISMATCH"] += 1
            print("MISMATCH (error %r)" % e, desc, flush=True)
            continue
        ok = (got.width, got.height) == (r["width"], r["height"]) and got.config == {1: got.config, 2: "ARGB_8888", 3: "RGBA_F16", 4: "RGB_565", 5: "RGBA_1010102"}[cfg]
        if ok:
            bpp = {"ARGB_8888": 4, "RGBA_F16": 8, "RGB_565": 2, "RGBA_1010102": 4}[got.config]
            a = np.ascontiguousarray(got.pixels[:, : got.width * bpp])
            b = np.ascontiguousarray(r["pixels"][:, : got.width * bpp])
            if got.config == "ARGB_8888":
                d = np.abs(a.astype(int) - b.astype(int))
                tol = 2 if sampled or alpha else 1
                ok = d.max() <= tol and (d == 0).mean() > (0.9 if alpha or sampled else 0.97)
                info = (int(d.max()), float((d == 0).mean()))
            elif got.config == "RGBA_F16":
                fa, fb = a.view(np.float16).astype(np.float32), b.view(np.float16).astype(np.float32)
                d = np.abs(fa - fb)
                ok = d.max() <= 2.5 / 255 and (d == 0).mean() > 0.9
                info = (float(d.max()), float((d == 0).mean()))
            elif got.config == "RGBA_1010102":
                ua, ub = a.view(np.uint32), b.view(np.uint32)
                dm = 0
                for sh in (0, 10, 20):
                    dm = max(dm, int(np.abs(((ua >> sh) & 0x3FF).astype(int) - ((ub >> sh) & 0x3FF).astype(int)).max()))
                ok = dm <= 8 and ((ua >> 30) == (ub >> 30)).mean() > 0.99
                info = (dm,)
            else:
                ua, ub = a.view(np.uint16), b.view(np.uint16)
                dm = max(int(np.abs(((ua >> 11) & 31).astype(int) - ((ub >> 11) & 31).astype(int)).max()),
                         int(np.abs(((ua >> 5) & 63).astype(int) - ((ub >> 5) & 63).astype(int)).max()),
                         int(np.abs((ua & 31).astype(int) - (ub & 31).astype(int)).max()))
                ok = dm <= 1
                info = (dm,)
        else:
            info = ("shape/config", got.width, got.height, got.config, r["width"], r["height"])
        if not ok and tf == 16 and (got.width, got.height) == (r["width"], r["height"]) and got.config in ("ARGB_8888", "RGBA_1010102"):
            if got.config == "ARGB_8888":
                da, rb, tol, dark = np.abs(a.astype(int) - b.astype(int)), b.astype(int), 2 if sampled or alpha else 1, 64
            else:
                ua, ub = a.view(np.uint32), b.view(np.uint32)
                da = np.stack([np.abs(((ua >> sh) & 0x3FF).astype(int) - ((ub >> sh) & 0x3FF).astype(int)) for sh in (0, 10, 20)])
                rb, tol, dark = np.stack([((ub >> sh) & 0x3FF).astype(int) for sh in (0, 10, 20)]), 8, 256
            far = da > tol
            if float(far.mean()) <= (1e-3 if sampled else 3e-5) and int(rb[far].max()) < dark:
                counts["MATCH_PQ_DARK"] += 1
                print("MATCH_PQ_DARK", info, "far share %.2e, brightest reference value there %d" % (float(far.mean()), int(rb[far].max())), flush=True)
                continue
        if ok:
            counts["MATCH"] += 1
        else:
            counts["MISMATCH"] += 1
            print("MISMATCH", info, desc, flush=True)
            if got.config == "ARGB_8888" and (got.width, got.height) == (r["width"], r["height"]):
                bad = (a.reshape(got.height, got.width, 4) != b.reshape(got.height, got.width, 4)).any(axis=2)
                ys, xs = np.nonzero(bad)
                grid = {}
                for gy in range(0, got.height, 256):
                    for gx in range(0, got.width, 256):
                        grid[(gx // 256, gy // 256)] = round(float(bad[gy:gy + 256, gx:gx + 256].mean()), 3)
                print("   bbox x %d..%d y %d..%d; mismatch share per 256-group: %s; channels differing: %s" % (
                    xs.min(), xs.max(), ys.min(), ys.max(), grid,
                    [int((a.reshape(got.height, got.width, 4)[..., c] != b.reshape(got.height, got.width, 4)[..., c]).sum()) for c in range(4)]), flush=True)
            if os.environ.get("SWEEP_DUMP"):
                open(os.path.join(ROOT, "gpurun_out", "sweep_case_%d.jxl" % k), "wb").write(data)
    print(counts)
    print("refused:", refused)


if __name__ == "__main__":
    main()
