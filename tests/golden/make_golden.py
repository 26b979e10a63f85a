"""Generates the committed golden vectors: small JXL inputs (encoded with the reference's own encoder settings) and the
pixels the REFERENCE produces for them (oracle/_ref = its unmodified JNI decode path + prebuilt libjxl 0.12.0/weaver).
Run where /root/reference exists:  python tests/golden/make_golden.py
Each <name>.npz holds: jxl (uint8 bytes), raw (DecodeJpegXlOneShot RGBA), and out_<cfg> = decodeSampled(cfg) pixels
(Bitmap bytes [h, stride]) for the colour configs listed."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402
from oracle import refjxl  # noqa: E402

GOLDEN = ["rgb_lossy_64", "rgb_lossy_64_epf3", "rgb_lossy_64_epf2_nogab", "rgb_lossy_256x200", "rgba_lossy_300x203",
          "rgb_lossy_2304x24", "rgba_lossless_128", "rgb_lossless_200x150", "natural_d1_e7", "natural_d4_e7"]
CFGS = {"rgba_lossless_128": (1, 2, 3, 4, 5), "rgb_lossless_200x150": (1, 2, 3, 4, 5), "rgb_lossy_64": (2,), "rgba_lossy_300x203": (2,)}


def main():
    for name in GOLDEN:
        data = cases.get(name)
        raw, meta = refjxl.decode_oneshot(data)
        arrs = dict(jxl=np.frombuffer(data, np.uint8), raw=raw)
        for cfg in CFGS.get(name, ()):
            r = refjxl.decode_sampled(data, cfg=cfg)
            arrs["out_%d" % cfg] = r["pixels"]
            arrs["cfg_%d" % cfg] = np.array([ord(c) for c in r["config"]], np.uint8)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
        print(name, len(data), raw.shape)


if __name__ == "__main__":
    main()
