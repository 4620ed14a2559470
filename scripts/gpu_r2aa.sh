#!/bin/bash
# round 2, GPU call AA: flat filter with parked call arguments and ordered subtractions
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "filter_the_super or shortcuts or image_sweep or update_list" 2>&1 | tail -2
timeout 200 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2aa.json 2> gpurun_out/r2aa.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2aa.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"], 4), "sweep", round(d["roofline"]["kernel_ms"], 4), "rebuild_ms (filter)", round(d["rebuild_ms"], 3), d["rebuild_kinds"])
PY
