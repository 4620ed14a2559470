#!/bin/bash
# round 2, GPU call S (2 GPUs): where does the in-process group test wait (B200COORD_TRACE)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -k "group" > gpurun_out/r2s_group.log 2>&1
echo "group rc=$?" >> gpurun_out/r2s_group.log
tail -40 gpurun_out/r2s_group.log | cut -c1-200
timeout 200 python -m pytest tests/test_gpu_plugin_dropin.py -m gpu -x -q -k "devices" 2>&1 | tail -2
