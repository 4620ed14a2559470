#!/bin/bash
# round 2, GPU call D (2 GPUs): full GPU test suite incl. the two-rank tests, 1- and 2-GPU bench with the pull-based exchange
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2d_gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -5 gpurun_out/r2d_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-other-configs > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err
tail -3 gpurun_out/r2d_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-other-configs --no-peer > gpurun_out/r2d_bench_n2_nopeer.json 2> gpurun_out/r2d_bench_n2_nopeer.err
python - <<'PY'
import json
for f in ("n1", "n2", "n2_nopeer"):
    try:
        d = json.loads(open("gpurun_out/r2d_bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.3f sweep %.3f e2e ms %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"]), d.get("parity_check"))
    except Exception as e:
        print(f, "failed", e)
PY
