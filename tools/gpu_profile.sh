#!/usr/bin/env bash
# Run on the GPU box (under gpurun): launch list of OUR kernels for one bench step, then a full capture of the top kernels.
# usage: tools/gpu_profile.sh <tag> [full-capture kernel regex]
set -u
TAG="${1:-x}"
FULL="${2:-}"
mkdir -p gpurun_out
KF='regex:^(k_|Device)'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_bench_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_bench_${TAG}.log | cut -c1-300
if [ -n "$FULL" ]; then
  # skip the warm-up pass (-s counted per matching kernel) and capture the timed step's launches
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$FULL" -c 12 -o gpurun_out/full_${TAG} -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_${TAG}.log 2>&1
  tail -2 gpurun_out/ncu_full_${TAG}.log | cut -c1-300
  ls -la gpurun_out/full_${TAG}.ncu-rep
fi
