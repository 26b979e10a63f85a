"""GPU diagnostic (not a test): decodes every golden / cached case through the C ABI and prints parity numbers against
the golden vectors, the CPU emulation of the same code and (if present) the reference; never asserts."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_lib  # noqa: E402
import hostemu_lib as H  # noqa: E402
import jxl_coder_b200 as J  # noqa: E402


def premul(a):
    a = a.copy()
    if a[..., 3].min() < 255:
        al = a[..., 3:4].astype(np.uint16)
        a[..., :3] = (a[..., :3].astype(np.uint16) * al // 255).astype(np.uint8)
    return a


def main():
    import torch
    print("cuda:", torch.cuda.is_available(), torch.cuda.get_device_name(0) if torch.cuda.is_available() else None, flush=True)
    print(J.load_library().jxlb_version())
    for name in golden_lib.names():
        g = golden_lib.load(name)
        try:
            t = time.time()
            bmp = J.JxlCoder.decode(g["jxl"], 2)
            dt = time.time() - t
            out = bmp.as_array()
            want = premul(g["raw"])
            d = np.abs(out.astype(int) - want.astype(int))
            e = H.Decoded(g["jxl"])
            emu = premul(e.render())
            e.close()
            de = np.abs(out.astype(int) - emu.astype(int))
            print("%-28s vs golden: exact %.4f max %d | vs cpu-emu: exact %.5f max %d | %.1f ms %s" % (
                name, (d == 0).mean(), d.max(), (de == 0).mean(), de.max(), dt * 1e3, J.last_batch_timings()), flush=True)
        except Exception as ex:
            print(name, "FAILED", repr(ex), flush=True)
            traceback.print_exc()


if __name__ == "__main__":
    main()
