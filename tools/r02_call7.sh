#!/bin/bash
# Round-2 GPU call 7: the GPU BVH builder (lrb_build_lbvh / EMBREE_MORTON): tests, build times, traversal speed on its trees.
set -u
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-600}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
T=900 run python -m pytest tests/test_gpu_builder.py -q -x
T=300 run python tools/builder_bench.py
T=300 run python bench.py --builder EMBREE_MORTON --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02c7_bench_kitchen_morton.err | tee gpurun_out/r02c7_bench_kitchen_morton.json | cut -c1-300
ls -la gpurun_out | tail -5
