#!/bin/bash
# Round-2 GPU call 16: the whole GPU suite, smoke() and the default bench line on the final commit.
set -u
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-600}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
T=1500 run python -m pytest tests -m gpu -q -x
T=300 run python __graft_entry__.py smoke
T=400 run python bench.py 2>gpurun_out/r02c16_bench.err | tee gpurun_out/r02c16_bench.json | cut -c1-300
T=300 run python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/r02c16_bench_ref.err | tee gpurun_out/r02c16_bench_ref.json | cut -c1-300
