#!/bin/bash
# round 2, GPU call A: parity of the image-mode sweep, A/B timings, one ncu capture of the new kernel
set -x
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2b_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
for v in 2; do
  B200COORD_IMG_VARIANT=$v timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2b_bench_v$v.json 2> gpurun_out/r2b_bench_v$v.err
done
B200COORD_NO_IMG_SWEEP=1 timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2b_bench_noimg.json 2> gpurun_out/r2b_bench_noimg.err
B200COORD_NO_FAR_SPLIT=1 timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2b_bench_nofar.json 2> gpurun_out/r2b_bench_nofar.err
B200COORD_NO_SUPERLIST=1 timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2b_bench_nosuper.json 2> gpurun_out/r2b_bench_nosuper.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_img -s 4 -c 1 -f -o gpurun_out/prof_img_r2b \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --frames 2 > gpurun_out/prof_img_r2b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2b.csv \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --frames 2 > gpurun_out/launches_r2b.log 2>&1
grep -h '"value"' gpurun_out/r2b_bench_*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['rebuild_ms'], d['e2e']['ms_per_step'], d['roofline']['frac'])
"
