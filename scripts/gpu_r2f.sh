#!/bin/bash
# round 2, GPU call F: cell-tile sweep after the range-table fix (short timeouts: a hang must not burn the budget)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -5 gpurun_out/r2f_pytest.log
for cfgname in "configs[0]" "configs[2]-NLIST" "configs[2]-NLISTCELLS"; do
  timeout 120 python bench.py --only-other "$cfgname" > "gpurun_out/r2f_only_$cfgname.json" 2> "gpurun_out/r2f_only_$cfgname.err"
  echo "rc=$?"; cat "gpurun_out/r2f_only_$cfgname.json"
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2f_cfg2nlist.csv python bench.py --only-other "configs[2]-NLIST" > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2f_cfg2cells.csv python bench.py --only-other "configs[2]-NLISTCELLS" > /dev/null 2>&1
B200COORD_PLUGIN_TIMERS=1 timeout 420 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_full.json 2> gpurun_out/r2f_bench_full.err
echo "full rc=$?"; tail -5 gpurun_out/r2f_bench_full.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2f_bench_full.json").read().strip().splitlines()[-1])
    print("typical", d["ms_per_step"], d["roofline"]["kernel_ms"], "e2e", d["e2e"]["ms_per_step"])
    for k, v in d["other_configs"].items():
        print(k, {q: v.get(q) for q in ("ms_per_step", "sweep_ms", "rebuild_ms", "roofline_frac", "cv_value")})
    print(json.dumps(d.get("cuda_baseline"), indent=1))
    print(json.dumps(d.get("e2e_plumed"), indent=1)[:1500])
except Exception as e:
    print("no full json", e)
PY
