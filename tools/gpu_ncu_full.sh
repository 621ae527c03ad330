#!/usr/bin/env bash
# Run on the GPU box: one `ncu --set full` capture per kernel regex (first timed-step launch of each), bench volume.
# usage: tools/gpu_ncu_full.sh <tag> <regex1> [regex2 ...]
TAG="$1"; shift
mkdir -p gpurun_out
for RX in "$@"; do
  NAME=$(echo "$RX" | tr -c 'A-Za-z0-9_' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$RX" -s 1 -c 1 -o gpurun_out/full_${TAG}_${NAME} -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_${TAG}_${NAME}.log 2>&1
  ls -la gpurun_out/full_${TAG}_${NAME}.ncu-rep 2>&1 | tail -1
done
