#!/bin/bash
# GPU-box visit: parity tests, AC lanes-per-warp sweep, overlapped bench.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
for L in 32 16 8 4; do
  echo "== AC lanes $L"; JXLB_AC_LANES=$L timeout 300 python tests/gpu_perf.py 64 4096 2 2>&1 | grep -E "^run|vs reference" | tee -a gpurun_out/ac_lanes.txt
done
echo "== bench"; timeout 900 python bench.py --steps 6 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.json
echo "== bench contexts=1"; timeout 900 python bench.py --steps 4 --warmup 3 --contexts 1 --callers 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c1.json
