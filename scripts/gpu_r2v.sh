#!/bin/bash
# round 2, GPU call V: the final code once more: whole suite, smoke, the driver's bench line
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
tail -3 gpurun_out/r2v_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2v_full.json 2> gpurun_out/r2v_full.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2v_full.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "sweep", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "ahead", (d["e2e"].get("frames_submitted_ahead") or {}).get("ms_per_step"))
e = d["e2e_plumed"]; print("e2e_plumed", e["ms_per_step"], "coupled", e["device_coupled"].get("ms_per_step"))
print("launches", d["gpu_launches"], "clocks", d["clocks"])
PY
