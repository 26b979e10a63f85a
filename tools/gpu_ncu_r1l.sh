#!/bin/bash
# ncu --set full of every kernel of the main chain, one 4096x4096 image per launch (tests/gpu_perf.py 1 image, 1 run).
set -u
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"LfGroupKernel|AcLaneKernel|LfFinalKernel|ReconRegionKernel|ReconLargeKernel|FilterColorFastKernel|FrameStatusKernel|BuildGroupBlocksKernel" -c 10 -f -o gpurun_out/prof_r1l_all python tests/gpu_perf.py 1 4096 1 > gpurun_out/prof_r1l_all.log 2>&1
tail -3 gpurun_out/prof_r1l_all.log
timeout 300 ncu -i gpurun_out/prof_r1l_all.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct > gpurun_out/prof_r1l_all.csv 2>&1
wc -l gpurun_out/prof_r1l_all.csv
