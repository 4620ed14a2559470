#!/bin/bash
# round 2, GPU call L: plugin host path after the addForces fix, device coupling, flat filter kernel
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_plugin_dropin.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -5 gpurun_out/r2l_pytest.log
for flat in 0 1; do
  B200COORD_FILTER_FLAT=$flat timeout 300 python bench.py --steps 40 --warmup 10 --quick --no-cpu-baseline > gpurun_out/r2l_flat$flat.json 2> gpurun_out/r2l_flat$flat.err
done
B200COORD_PLUGIN_TIMERS=1 timeout 420 python bench.py --gpus 1 --steps 20 --warmup 5 --no-regimes --no-other-configs --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
grep "plugin timers" gpurun_out/r2l_bench.err
python - <<'PY'
import json
for flat in (0, 1):
    try:
        d = json.loads(open("gpurun_out/r2l_flat%d.json" % flat).read().strip().splitlines()[-1])
        print("flat", flat, "ms/step", d["ms_per_step"], "sweep", d["roofline"]["kernel_ms"], "rebuild_ms", d["rebuild_ms"], d["rebuild_kinds"])
    except Exception as e:
        print("flat", flat, "failed", e)
d = json.loads(open("gpurun_out/r2l_bench.json").read().strip().splitlines()[-1])
print("typical", d["ms_per_step"], d["roofline"]["kernel_ms"], "e2e", d["e2e"]["ms_per_step"])
e = d.get("e2e_plumed") or {}
print({k: v for k, v in e.items() if k not in ("plumed_timers", "device_coupled")})
for l in e.get("plumed_timers", []): print(l)
print("coupled", e.get("device_coupled"))
print(json.dumps(d.get("cuda_baseline"), indent=1))
PY
# the filter rebuild under ncu (one launch of the fill pass), both kernels
for flat in 0 1; do
  B200COORD_FILTER_FLAT=$flat timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_nl_filter -s 3 -c 1 -f -o gpurun_out/prof_filter_r2l_flat$flat python bench.py --steps 10 --warmup 3 --quick --no-cpu-baseline --frames 2 > gpurun_out/prof_filter_r2l_flat$flat.log 2>&1
  tail -2 gpurun_out/prof_filter_r2l_flat$flat.log
done
