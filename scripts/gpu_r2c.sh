#!/bin/bash
# round 2, GPU call C: the new bench.py end to end, the reference arm at 100k atoms, ncu of the sweep with L2 prefetch
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "image or far_parts or super_list or smoke or update_list" > gpurun_out/r2c_pytest.log 2>&1
tail -3 gpurun_out/r2c_pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_full.json 2> gpurun_out/r2c_bench_full.err ) 2> gpurun_out/r2c_bench_full.time
tail -5 gpurun_out/r2c_bench_full.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_img -s 4 -c 1 -f -o gpurun_out/prof_img_r2c \
  python bench.py --steps 10 --warmup 3 --quick > gpurun_out/prof_img_r2c.log 2>&1
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err ) 2> gpurun_out/r2c_bench_ref.time
cat gpurun_out/r2c_bench_full.time gpurun_out/r2c_bench_ref.time
head -c 1500 gpurun_out/r2c_bench_full.json
