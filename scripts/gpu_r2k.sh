#!/bin/bash
# round 2, GPU call K: parallel host path of the plugin, plumed benchmark A/B harness, whole suite
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -5 gpurun_out/r2k_pytest.log
B200COORD_PLUGIN_TIMERS=1 timeout 420 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
tail -6 gpurun_out/r2k_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2k_bench.json").read().strip().splitlines()[-1])
print("typical", d["ms_per_step"], d["roofline"]["kernel_ms"], "e2e", d["e2e"]["ms_per_step"])
print(json.dumps(d.get("e2e_plumed"), indent=1)[:2500])
print(json.dumps(d.get("cuda_baseline"), indent=1))
print(d["other_configs"]["configs[3]"])
PY
timeout 300 bash scripts/plumed_benchmark_ab.sh 20000 200 > gpurun_out/r2k_plumed_benchmark_ab.txt 2>&1
grep "BENCH:" gpurun_out/r2k_plumed_benchmark_ab.txt | head -30
