#!/usr/bin/env bash
# Run on the GPU box: per-kernel device times (ncu launch list) of the kernels matching a regex, for env variants.
# usage: tools/gpu_kernel_times.sh <tag> <kernel regex> "<ENV1=.. ENV2=..>" ["<variant 2>" ...]
TAG="$1"; RX="$2"; shift 2
mkdir -p gpurun_out
i=0
for V in "$@"; do
  i=$((i+1))
  env $V CKL_CHUNKS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:$RX" --csv \
    --log-file gpurun_out/kt_${TAG}_$i.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/kt_${TAG}_$i.log 2>&1
  echo "== $V"
  python tools/launch_summary.py gpurun_out/kt_${TAG}_$i.csv | tail -n +2
done
