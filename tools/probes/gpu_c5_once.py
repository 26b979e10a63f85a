"""GPU probe (not a test): configs[4] with the current JXLB_ANIM_PREFETCH."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import jxl_coder_b200 as J  # noqa: E402
import gpu_configs  # noqa: E402

gpu_configs.c5()
print(J.last_batch_timings())
