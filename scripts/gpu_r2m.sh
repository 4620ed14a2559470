#!/bin/bash
# round 2, GPU call M: where does an occasional long rebuild come from; filter captures; a driver-style full run
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 300 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2m_quick$i.json 2> gpurun_out/r2m_quick$i.err
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2m_full.json 2> gpurun_out/r2m_full.err
python - <<'PY'
import json
for f in ("quick1", "quick2", "quick3", "full"):
    try:
        d = json.loads(open("gpurun_out/r2m_%s.json" % f).read().strip().splitlines()[-1])
        t = d["regimes"]["typical"]
        print(f, "ms/step", round(d["ms_per_step"], 4), "sweep", round(d["roofline"]["kernel_ms"], 4), "rebuild avg/max", round(t["rebuild_ms"], 3), t.get("rebuild_ms_max"), "e2e", round(d["e2e"]["ms_per_step"], 3))
        if f == "full":
            for k, v in d["regimes"].items(): print("  ", k, round(v["ms_per_step"], 4), round(v["rebuild_ms"], 3), v.get("rebuild_ms_max"))
            print("   sustained", d["sustained"])
            e = d["e2e_plumed"]; print("   e2e_plumed", e["ms_per_step"], "coupled", e["device_coupled"].get("ms_per_step"))
    except Exception as e:
        print(f, "failed", repr(e))
PY
for flat in 0 1; do
  B200COORD_FILTER_FLAT=$flat timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_nl_filter -s 2 -c 1 -f -o gpurun_out/prof_filter_r2m_flat$flat python bench.py --steps 12 --warmup 3 --quick > gpurun_out/prof_filter_r2m_flat$flat.log 2>&1
  tail -2 gpurun_out/prof_filter_r2m_flat$flat.log
done
