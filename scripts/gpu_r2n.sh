#!/bin/bash
# round 2, GPU call N: flat filter after the instruction diet (2 or 3 resident blocks), parity first
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -4 gpurun_out/r2n_pytest.log
B200COORD_FILTER_MINB=3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "filter or shortcuts or update_list or one_million" > gpurun_out/r2n_pytest_minb3.log 2>&1
tail -2 gpurun_out/r2n_pytest_minb3.log
for v in "FLAT=0" "MINB=2" "MINB=3"; do
  env B200COORD_FILTER_$v timeout 300 python bench.py --steps 40 --warmup 11 --quick > gpurun_out/r2n_$v.json 2> gpurun_out/r2n_$v.err
done
python - <<'PY'
import json
for v in ("FLAT=0", "MINB=2", "MINB=3"):
    try:
        d = json.loads(open("gpurun_out/r2n_%s.json" % v).read().strip().splitlines()[-1])
        t = d["regimes"]["typical"]
        print(v, "ms/step", round(d["ms_per_step"], 4), "sweep", round(d["roofline"]["kernel_ms"], 4), "rebuild", round(t["rebuild_ms"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3))
    except Exception as e:
        print(v, "failed", repr(e))
PY
for v in 2 3; do
  B200COORD_FILTER_MINB=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_nl_filter -s 2 -c 1 -f -o gpurun_out/prof_filter_r2n_minb$v python bench.py --steps 12 --warmup 3 --quick > gpurun_out/prof_filter_r2n_minb$v.log 2>&1
  tail -1 gpurun_out/prof_filter_r2n_minb$v.log
done
