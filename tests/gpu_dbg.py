"""GPU debug helper (not a test): decode a few inputs and print status / message."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import jxl_coder_b200 as J
import cases
names = sys.argv[1:] or ["rgb_lossy_64", "rgb_lossy_256x200", "natural_512_d1", "rgb_lossy_2304x24"]
for n in names:
    data = open(n, "rb").read() if os.path.exists(n) else cases.get(n)
    try:
        out = J.JxlCoder.decode(data, 2)
        print(n, "ok", out.as_array().shape, flush=True)
    except Exception as e:
        print(n, "FAIL", repr(e), flush=True)
