#!/bin/bash
# round 2, GPU call Q (2 GPUs): the multi-GPU paths with the final rebuild kernels; frames submitted ahead
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_plugin_dropin.py tests/test_gpu_parity.py -m gpu -x -q -k "multirank or rank or devices or group or submitted_ahead or distributed" > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -4 gpurun_out/r2q_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2q_n2.json 2> gpurun_out/r2q_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2q_n2.json").read().strip().splitlines()[-1])
print("N=2 value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "parity", d.get("parity_check"))
PY
tail -3 gpurun_out/r2q_n2.err
