#!/usr/bin/env bash
# GPU tuning aid: compress time of the bench volume at K chunks under the scheduling knobs
K="${1:-4}"
for mc in 8 32; do for pr in 0 1; do for gm in 1 8; do
  echo "== maxconn $mc prio $pr gridmult $gm"
  CUDA_DEVICE_MAX_CONNECTIONS=$mc CKL_PRIO=$pr CKL_GRID_MULT=$gm timeout 120 python tools/chunk_tune.py 1024,1024,1024 $K 2>&1 | tail -1
  CUDA_DEVICE_MAX_CONNECTIONS=$mc CKL_PRIO=$pr CKL_GRID_MULT=$gm timeout 120 python tools/timeline.py 1024,1024,1024 $K 2> gpurun_out/tl_${K}_${mc}_${pr}_${gm}.txt
done; done; done
