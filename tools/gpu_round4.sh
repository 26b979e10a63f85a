#!/bin/bash
# Round-end style visit: parity tests, smoke, bench (both arms), ncu launch list of one bench-like step.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke.txt
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
echo "== bench"; timeout 1200 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.json
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tests/gpu_perf.py 8 4096 1 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
