#!/bin/bash
# round 2, GPU call E: cell-tile sweep (parity + configs[0]/[2] timings), A/B of the issue-side placement, PLUMED timers
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -5 gpurun_out/r2e_pytest.log
for v in 2 3 2 3; do
  B200COORD_IMG_VARIANT=$v timeout 300 python bench.py --steps 50 --warmup 10 --quick > gpurun_out/r2e_bench_v$v.json 2>> gpurun_out/r2e_bench_v.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2e_bench_v$v.json').read().strip().splitlines()[-1]); print('variant $v', d['ms_per_step'], d['roofline']['kernel_ms'], d['sustained']['ms_per_step'], d['sustained']['sweep_ms'])"
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_bench_full.json 2> gpurun_out/r2e_bench_full.err
tail -3 gpurun_out/r2e_bench_full.err
B200COORD_NO_TILE_SWEEP=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-regimes --no-plumed-e2e > gpurun_out/r2e_bench_notile.json 2> gpurun_out/r2e_bench_notile.err
python - <<'PY'
import json
for f in ("full", "notile"):
    d = json.loads(open("gpurun_out/r2e_bench_%s.json" % f).read().strip().splitlines()[-1])
    for k, v in d["other_configs"].items():
        print(f, k, {q: v.get(q) for q in ("ms_per_step", "sweep_ms", "rebuild_ms", "roofline_frac", "cv_value")})
    if "e2e_plumed" in d:
        print(json.dumps(d["e2e_plumed"], indent=1)[:3000])
PY
cp /tmp/bench_plumed_e2e.log gpurun_out/r2e_plumed_e2e.log 2>/dev/null
