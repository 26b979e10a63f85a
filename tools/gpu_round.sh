#!/bin/bash
# One GPU visit (run under gpurun).  Usage: tools/gpu_round.sh <tag> <stage> [<stage> ...]
#   tests              pytest -m gpu
#   smoke              __graft_entry__.smoke()
#   bench / benchref   bench.py (default flags) / bench.py --impl reference
#   benchq             short bench (--steps 6 --warmup 3, no CPU baseline)
#   perf[:N]           tools/probes/gpu_perf.py N 4096 3   (stage times of a prepared batch, default N=16)
#   configs            tools/probes/gpu_configs.py (the other BASELINE configs through the public API)
#   launches           ncu launch list (gpu__time_duration.sum) of an 8-image step
#   ncu:<regex>        ncu --set full of the kernels matching <regex>, one 4096x4096 image, + raw-page CSV + stall summary
#   sweep[:N]          randomised parity sweep against the reference (tools/probes/gpu_sweep.py), N cases x 2 seeds
# Everything lands in gpurun_out/<tag>_*.
set -u
tag=$1; shift
mkdir -p gpurun_out
for stage in "$@"; do
  name=${stage%%:*}; arg=""; [[ "$stage" == *:* ]] && arg=${stage#*:}
  echo "== $stage"
  case $name in
    tests) timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${tag}_pytest_gpu.txt ;;
    smoke) timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.txt ;;
    bench) timeout 1500 python bench.py 2>gpurun_out/${tag}_bench.err | tail -1 | tee gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err ;;
    benchq) timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>gpurun_out/${tag}_benchq.err | tail -1 | tee gpurun_out/${tag}_benchq.json; tail -3 gpurun_out/${tag}_benchq.err ;;
    benchref) timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${tag}_bench_ref.json ;;
    perf) timeout 600 python tools/probes/gpu_perf.py ${arg:-16} 4096 3 2>&1 | tee gpurun_out/${tag}_perf.txt ;;
    configs) timeout 900 python tools/probes/gpu_configs.py 2>&1 | tail -14 | tee gpurun_out/${tag}_configs.txt ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv python tools/probes/gpu_perf.py 8 4096 1 > gpurun_out/${tag}_launches.log 2>&1
      python tools/launch_summary.py gpurun_out/${tag}_launches.csv | tee gpurun_out/${tag}_launch_summary.txt ;;
    ncu)
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$arg" -c 12 -f -o gpurun_out/${tag}_prof python tools/probes/gpu_perf.py 1 4096 1 > gpurun_out/${tag}_prof.log 2>&1
      tail -2 gpurun_out/${tag}_prof.log
      timeout 300 ncu -i gpurun_out/${tag}_prof.ncu-rep --page raw --csv > gpurun_out/${tag}_prof_raw.csv 2>&1
      python tools/ncu_summary.py gpurun_out/${tag}_prof_raw.csv | tee gpurun_out/${tag}_ncu_summary.txt ;;
    sweep) for s in 11 12; do timeout 1200 python tools/probes/gpu_sweep.py $s ${arg:-125} 2>&1 | tail -25; done | tee gpurun_out/${tag}_sweep.txt ;;
    *) echo "unknown stage $stage" ;;
  esac
done
