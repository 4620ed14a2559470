#!/bin/bash
# round 2, GPU call P: frames submitted ahead (submit / collect): parity and the overlap it buys
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "submitted_ahead or device_resident or engine" > gpurun_out/r2p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -4 gpurun_out/r2p_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2p_quick.json 2> gpurun_out/r2p_quick.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2p_quick.json").read().strip().splitlines()[-1])
print("value ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "ahead", d["e2e"].get("frames_submitted_ahead"), "cv", d["cv_value"])
PY
tail -3 gpurun_out/r2p_quick.err
