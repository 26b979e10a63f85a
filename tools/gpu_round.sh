#!/bin/bash
# One GPU-box visit: parity tests, stage probe, bench (both arms), ncu launch list.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== perf probe"; timeout 600 python tests/gpu_perf.py 16 4096 3 2>&1 | tee gpurun_out/perf16.txt
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref.json
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tests/gpu_perf.py 8 4096 1 > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/ncu_launches.log
