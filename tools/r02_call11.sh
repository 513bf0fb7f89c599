#!/bin/bash
# Round-2 GPU call 11: the pipelined plugin sequence (tests + bench line), whole GPU suite on the new C ABI.
set -u
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-600}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
T=600 run python -m pytest tests/test_gpu_pipeline.py -q -x
T=400 run python bench.py --steps 10 --warmup 3 2>gpurun_out/r02c11_bench_kitchen.err | tee gpurun_out/r02c11_bench_kitchen.json | cut -c1-300
T=1500 run python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_pipeline.py
