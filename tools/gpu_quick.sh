#!/bin/bash
# Quick GPU check: parity tests + stage probe.  $1 = images (default 16)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
timeout 600 python tests/gpu_perf.py ${1:-16} 4096 3 2>&1 | tee gpurun_out/perf.txt
