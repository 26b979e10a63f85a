#!/bin/bash
# Quick GPU visit: parity tests + bench (ours only).  Extra args are passed to bench.py.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -2 | tee gpurun_out/bench_quick.json
