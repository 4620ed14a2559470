#!/bin/bash
# round 2, GPU call L: plugin host path after the addForces fix, device coupling
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_plugin_dropin.py -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -5 gpurun_out/r2l_pytest.log
B200COORD_PLUGIN_TIMERS=1 timeout 420 python bench.py --gpus 1 --steps 20 --warmup 5 --no-regimes --no-other-configs --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
grep "plugin timers" gpurun_out/r2l_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2l_bench.json").read().strip().splitlines()[-1])
print("typical", d["ms_per_step"], d["roofline"]["kernel_ms"], "e2e", d["e2e"]["ms_per_step"])
e = d.get("e2e_plumed") or {}
print({k: v for k, v in e.items() if k not in ("plumed_timers", "device_coupled")})
print("coupled", e.get("device_coupled"))
for l in e.get("plumed_timers", []): print(l)
print(json.dumps(d.get("cuda_baseline"), indent=1))
PY
# the filter rebuild under ncu (one launch of the fill pass)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_nl_filter -s 3 -c 1 -f -o gpurun_out/prof_filter_r2l python bench.py --steps 10 --warmup 3 --quick --no-cpu-baseline --frames 2 > gpurun_out/prof_filter_r2l.log 2>&1
tail -3 gpurun_out/prof_filter_r2l.log
