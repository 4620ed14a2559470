#!/bin/bash
# round 2, GPU call J (2 GPUs): group-in-one-process API, GPU_DEVICES keyword, full 1M oracle comparison, whole suite
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_plugin_dropin.py -x -q -m gpu -k "group or gpu_devices or top_level" > gpurun_out/r2j_group.log 2>&1
echo "rc=$?" >> gpurun_out/r2j_group.log
tail -15 gpurun_out/r2j_group.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -5 gpurun_out/r2j_pytest.log
