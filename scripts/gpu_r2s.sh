#!/bin/bash
# round 2, GPU call S (2 GPUs): where does the in-process group test wait (B200COORD_TRACE)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B200COORD_TRACE=1 timeout 75 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -s -k "group" > gpurun_out/r2s_group.log 2>&1
echo "group rc=$?" >> gpurun_out/r2s_group.log
tail -40 gpurun_out/r2s_group.log | cut -c1-200
