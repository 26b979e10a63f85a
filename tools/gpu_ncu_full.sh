#!/bin/bash
# ncu --set full capture (with source) of the kernels matching $1 (regex) while tests/gpu_perf.py decodes $2 images once.
set -u
mkdir -p gpurun_out
PAT=${1:-AcLaneKernel}
N=${2:-8}
NAME=${3:-prof}
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$PAT" -c ${4:-2} -f -o gpurun_out/$NAME python tests/gpu_perf.py $N 4096 1 > gpurun_out/$NAME.log 2>&1
tail -4 gpurun_out/$NAME.log
ls -la gpurun_out/$NAME.ncu-rep
