"""A fixed small instance of the randomised parity sweep (tools/probes/cpu_sweep.py): encoder settings drawn at random,
the reference's decode against the host build of the kernel code.  Lossless: bit-exact.  Lossy: one 8-bit step."""
import importlib.util
import os

import numpy as np

import hostemu_lib as H

_spec = importlib.util.spec_from_file_location("cpu_sweep", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                       "tools", "probes", "cpu_sweep.py"))
S = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(S)


def test_random_encoder_settings(ref):
    rng = np.random.default_rng(4242)
    bad, done = [], 0
    for k in range(40):
        img, d = S.make_case(rng, k, 4242)
        if d["w"] * d["h"] > 500 * 400:  # keep the CPU suite short
            continue
        res, detail = S.run_case(ref, H, img, d)
        done += res == "ok"
        if res == "bad":
            bad.append((detail, d))
    assert not bad, bad
    assert done >= 15
