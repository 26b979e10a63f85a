#!/bin/bash
# GPU-box visit: parity tests, bench (both arms), ncu launch list, full ncu of recon/filter/AC kernels.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref.json
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tests/gpu_perf.py 8 4096 1 > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/ncu_launches.log
echo "== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"ReconRegionKernel|FilterColorKernel|AcLaneKernel|LfGroupKernel" -c 5 -f -o gpurun_out/prof_r1b python tests/gpu_perf.py 4 4096 1 > gpurun_out/prof_r1b.log 2>&1
tail -4 gpurun_out/prof_r1b.log
ls -la gpurun_out/
