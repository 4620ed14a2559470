#!/bin/bash
# round 2, GPU call X: the cell-scan list builder at 64 registers / 4 blocks per SM vs 80 registers / 3 blocks
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B200COORD_ROWS_MINB=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "filter_the_super or shortcuts or image_sweep" 2>&1 | tail -2
for mb in 3 2; do
  B200COORD_ROWS_MINB=$mb B200COORD_NO_SUPERLIST=1 timeout 200 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2x_minb$mb.json 2> gpurun_out/r2x_minb$mb.err
done
python - <<'PY'
import json
for mb in (3, 2):
    try:
        d = json.loads(open("gpurun_out/r2x_minb%d.json" % mb).read().strip().splitlines()[-1])
        print("rows minb", mb, "ms/step", round(d["ms_per_step"], 4), "rebuild_ms (cell scan)", round(d["rebuild_ms"], 3), d["rebuild_kinds"])
    except Exception as e:
        print(mb, "failed", repr(e))
PY
