#!/bin/bash
# Round-end style visit (r1j): parity tests, smoke, bench (both arms), ncu launch list of a bench-like step, ncu --set full of
# the post-decode kernels on configs[3].
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke.txt
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
echo "== bench"; timeout 1200 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.json
echo "== configs"; timeout 600 python tests/gpu_configs.py 2>&1 | tail -12 | tee gpurun_out/configs.txt
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tests/gpu_perf.py 8 4096 1 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
echo "== ncu full: post-decode kernels on configs[3]"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"Resize|ColorMatrix|PackKernel" -c 8 -f -o gpurun_out/prof_r1j_post python tests/gpu_c4_once.py > gpurun_out/prof_r1j_post.log 2>&1
tail -2 gpurun_out/prof_r1j_post.log
timeout 300 ncu -i gpurun_out/prof_r1j_post.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed > gpurun_out/prof_r1j_post.csv 2>&1
tail -3 gpurun_out/prof_r1j_post.csv
