#!/bin/bash
# round 2, GPU call I: full GPU suite, the driver's bench invocation, ncu captures for profiles/
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -4 gpurun_out/r2i_pytest.log
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err ) 2> gpurun_out/r2i_bench.time
cat gpurun_out/r2i_bench.time; tail -3 gpurun_out/r2i_bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_sweep_img|k_nl_filter" -s 6 -c 3 -f -o gpurun_out/prof_r2i_sweep python bench.py --steps 12 --warmup 3 --quick > gpurun_out/prof_r2i_sweep.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sweep_tile -s 4 -c 2 -f -o gpurun_out/prof_r2i_tile python bench.py --only-other "configs[2]-NLISTCELLS" > gpurun_out/prof_r2i_tile.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_r2i.csv python bench.py --steps 20 --warmup 5 --quick > /dev/null 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2i_bench.json").read().strip().splitlines()[-1])
print("typical", d["ms_per_step"], d["value"], "sweep", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("copies_alone_ms_per_step"))
print({k: (v["ms_per_step"], v["sweep_ms"], v["rebuild_ms"]) for k, v in d["regimes"].items()})
print("sustained", d["sustained"])
for k, v in d["other_configs"].items():
    print(k, {q: v.get(q) for q in ("ms_per_step", "sweep_ms", "rebuild_ms", "roofline_frac")})
print(d.get("cuda_baseline")); print(d.get("e2e_plumed", {}).get("ms_per_step")); print(d.get("cpu_baseline"))
PY
