#!/bin/bash
# Parity tests (incl. rescale), smoke, configs probe.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke.txt
echo "== configs"; timeout 600 python tests/gpu_configs.py 2>&1 | tail -12 | tee gpurun_out/configs.txt
