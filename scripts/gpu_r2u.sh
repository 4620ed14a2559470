#!/bin/bash
# round 2, GPU call U (8 GPUs): the driver's scaling invocation at N=8 with the final code; group tests incl. the injected failure
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2u_n8.json 2> gpurun_out/r2u_n8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2u_n8.json").read().strip().splitlines()[-1])
print("N=8 value", d["value"], "ms/step", d["ms_per_step"], "sweep", d["roofline"]["kernel_ms"], "rebuild", d["rebuild_ms"], "e2e", d["e2e"]["ms_per_step"], "copies", d["e2e"]["copies_alone_ms_per_step"], "parity", d.get("parity_check"))
print("configs[4]", (d.get("other_configs") or {}).get("configs[4]"))
PY
tail -2 gpurun_out/r2u_n8.err
timeout 150 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -k "group" > gpurun_out/r2u_group.log 2>&1
echo "group rc=$?" >> gpurun_out/r2u_group.log
tail -3 gpurun_out/r2u_group.log
