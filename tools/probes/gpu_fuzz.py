"""GPU robustness probe (run by tests/test_gpu_parity.py::test_corrupt_inputs_never_hang in a subprocess with a timeout):
mutated and truncated files through jxlb_decode_batch.  Every request must come back as a picture or as an error."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import jxl_coder_b200 as J  # noqa: E402


def main():
    import cases
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    rng = np.random.default_rng(seed)
    names = ["rgb_lossy_256x200", "rgba_lossless_128", "natural_512_d1", "rgba_lossy_300x203", "rgb_lossy_2304x24"]
    srcs = [cases.get(n) for n in names]
    p = os.path.join(ROOT, "tests", "_cache", "rgb_lossy_1024x768.jxl")
    if os.path.exists(p):
        srcs.append(open(p, "rb").read())
    p = os.path.join(ROOT, "tests", "_cache", "rgba_lossy_sq_320x264_s41_a1.jxl")
    if os.path.exists(p):
        srcs.append(open(p, "rb").read())
    batch = []
    for d in srcs:
        for k in range(24):
            b = bytearray(d)
            mode = k % 4
            if mode == 0:
                for _ in range(1 + k // 8):
                    b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            elif mode == 1:
                b = b[: int(rng.integers(10, len(b)))]
            elif mode == 2:
                b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
            else:
                b[int(rng.integers(0, min(len(b), 200)))] = int(rng.integers(0, 256))
            batch.append(bytes(b))
    ok = err = 0
    for i in range(0, len(batch), 16):
        res = J.decode_batch(batch[i:i + 16], config=2, raise_on_error=False)
        for r in res:
            if isinstance(r, J.Bitmap):
                ok += 1
            else:
                err += 1
    # the library must still decode a good file afterwards
    good = J.JxlCoder.decode(srcs[0], 2)
    assert good.width == 256
    print("fuzz ok=%d err=%d" % (ok, err))


if __name__ == "__main__":
    main()
