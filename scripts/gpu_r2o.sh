#!/bin/bash
# round 2, GPU call O: whole suite, driver-style bench, launch list
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -4 gpurun_out/r2o_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2o_full.json 2> gpurun_out/r2o_full.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2o_full.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "sweep", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"])
for k, v in d["regimes"].items(): print("  ", k, round(v["ms_per_step"], 4), "rebuild", round(v["rebuild_ms"], 3))
print("  sustained", d["sustained"]["ms_per_step"])
e = d["e2e_plumed"]; print("  e2e_plumed", e["ms_per_step"], "coupled", e["device_coupled"].get("ms_per_step"))
print("  other", json.dumps(d["other_configs"])[:1500])
print("  cuda", json.dumps(d["cuda_baseline"])[:800])
print("  cpu", d.get("cpu_baseline"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2o_launches.csv python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2o_launches.log 2>&1
tail -1 gpurun_out/r2o_launches.log | cut -c1-200
