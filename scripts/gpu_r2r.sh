#!/bin/bash
# round 2, GPU call R (2 GPUs): the in-process group after the collective_done rule (short leash: it hung before)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -k "group" > gpurun_out/r2r_pytest_group.log 2>&1
echo "group rc=$?" >> gpurun_out/r2r_pytest_group.log
tail -3 gpurun_out/r2r_pytest_group.log
timeout 240 python -m pytest tests/test_gpu_plugin_dropin.py -m gpu -x -q -k "devices" > gpurun_out/r2r_pytest_devices.log 2>&1
echo "devices rc=$?" >> gpurun_out/r2r_pytest_devices.log
tail -3 gpurun_out/r2r_pytest_devices.log
timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -k "distributed" > gpurun_out/r2r_pytest_dist.log 2>&1
echo "dist rc=$?" >> gpurun_out/r2r_pytest_dist.log
tail -3 gpurun_out/r2r_pytest_dist.log
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "submitted_ahead" > gpurun_out/r2r_pytest_ahead.log 2>&1
tail -2 gpurun_out/r2r_pytest_ahead.log
