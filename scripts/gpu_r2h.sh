#!/bin/bash
# round 2, GPU call H (8 GPUs): multi-rank parity on 2 and 8 ranks, the driver's scaling invocation at N=8 and N=4
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2h_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu -v > gpurun_out/r2h_multirank.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_multirank.log
tail -8 gpurun_out/r2h_multirank.log
run() {  # N, tag, extra flags
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus $1 --steps 20 --warmup 5 $3 > gpurun_out/r2h_bench_$2.json 2> gpurun_out/r2h_bench_$2.err
  echo "$2 rc=$?"; tail -2 gpurun_out/r2h_bench_$2.err
}
run 8 n8 ""
run 8 n8_nopeer "--no-peer --no-other-configs"
run 4 n4 "--no-other-configs"
timeout 300 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err
python - <<'PY'
import json
for f in ("n1", "n4", "n8", "n8_nopeer"):
    try:
        d = json.loads(open("gpurun_out/r2h_bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.3f sweep %.3f e2e %.4g ms %.3f numa %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["numa_node_of_rank0"]), d.get("parity_check"))
        if "other_configs" in d:
            print("   ", {k: (v.get("ms_per_step"), v.get("value")) for k, v in d["other_configs"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
