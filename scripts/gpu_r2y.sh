#!/bin/bash
# round 2, GPU call Y: how far ahead should the image sweep prefetch its entry stream into L2
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B200COORD_IMG_PREFETCH=4 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "image_sweep or scatter" 2>&1 | tail -2
for pf in 2 3 4; do
  B200COORD_IMG_PREFETCH=$pf timeout 200 python bench.py --steps 40 --warmup 11 --quick > gpurun_out/r2y_pf$pf.json 2> gpurun_out/r2y_pf$pf.err
done
python - <<'PY'
import json
for pf in (2, 3, 4):
    try:
        d = json.loads(open("gpurun_out/r2y_pf%d.json" % pf).read().strip().splitlines()[-1])
        print("prefetch", pf, "ms/step", round(d["ms_per_step"], 4), "sweep", round(d["roofline"]["kernel_ms"], 4), "sustained", round(d["sustained"]["ms_per_step"], 4), round(d["sustained"]["sweep_ms"], 4))
    except Exception as e:
        print(pf, "failed", repr(e))
PY
