#!/bin/bash
# ncu --set full capture with source of kernels matching $1 while gpu_perf decodes $2 images; report name $3
set -u
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$1" -c ${4:-3} -f -o gpurun_out/$3 python tests/gpu_perf.py ${2:-2} 4096 1 > gpurun_out/$3.log 2>&1
tail -2 gpurun_out/$3.log
