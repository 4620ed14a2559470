#!/bin/bash
# round 2, GPU call G: scatter path for short GROUPB rows, tile sweep with whole cells per block, device pair export,
# two-level finalize, cheap fall-back launch
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -5 gpurun_out/r2g_pytest.log
for cfgname in "configs[0]" "configs[2]-NLIST" "configs[2]-NLISTCELLS"; do
  timeout 120 python bench.py --only-other "$cfgname" > "gpurun_out/r2g_only_$cfgname.json" 2> "gpurun_out/r2g_only_$cfgname.err"
  echo "rc=$?"; python -c "
import json,sys
d=json.load(open('gpurun_out/r2g_only_$cfgname.json'))
print({k:{q:v.get(q) for q in ('ms_per_step','sweep_ms','rebuild_ms','cv_value')} for k,v in d.items()})"
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2g_cfg2nlist.csv python bench.py --only-other "configs[2]-NLIST" > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2g_cfg2cells.csv python bench.py --only-other "configs[2]-NLISTCELLS" > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2g_cfg0.csv python bench.py --only-other "configs[0]" > /dev/null 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2g_bench_quick.json 2> gpurun_out/r2g_bench_quick.err
python -c "
import json; d=json.loads(open('gpurun_out/r2g_bench_quick.json').read().strip().splitlines()[-1]); print('quick', d['ms_per_step'], d['roofline']['kernel_ms'], d['sustained']['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2g.csv python bench.py --steps 10 --warmup 3 --quick > /dev/null 2>&1
