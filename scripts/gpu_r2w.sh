#!/bin/bash
# round 2, GPU call W: quick check after the last library rebuild (copy streams synchronised in destroy)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plugin_dropin.py -m gpu -x -q -k "submitted_ahead or image_sweep or engine_arrays or host_fast" 2>&1 | tail -2
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
