#!/bin/bash
# round 2, GPU call Z: the final binary: whole suite, smoke, the driver's bench line
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
tail -3 gpurun_out/r2z_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_full.json 2> gpurun_out/r2z_full.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2z_full.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "sweep", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"])
for k, v in d["regimes"].items(): print("  ", k, round(v["ms_per_step"], 4), "rebuild", round(v["rebuild_ms"], 3))
o = d["other_configs"]
print("  c0", o["configs[0]"]["ms_per_step"], "c2", o["configs[2] NLIST"]["ms_per_step"], o["configs[2] NLISTCELLS"]["ms_per_step"], "c3", o["configs[3]"]["ms_per_step"], "c4", o["configs[4]"]["ms_per_step"])
PY
