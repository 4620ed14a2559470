#!/bin/bash
# round 2, GPU call T (4 GPUs): the driver's scaling invocation at N=4 with the final code
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2t_n4.json 2> gpurun_out/r2t_n4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2t_n4.json").read().strip().splitlines()[-1])
print("N=4 value", d["value"], "ms/step", d["ms_per_step"], "sweep", d["roofline"]["kernel_ms"], "rebuild", d["rebuild_ms"], "e2e", d["e2e"]["ms_per_step"], "parity", d.get("parity_check"))
print("configs[4]", (d.get("other_configs") or {}).get("configs[4]"))
PY
tail -3 gpurun_out/r2t_n4.err
