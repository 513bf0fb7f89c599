#!/bin/bash
# Round-2 GPU call 19: lrb_bvh_build_scene (device-side build + re-layout) -- its GPU tests, the GPU builder tests (the
# refactored lrb_build_bvh), stage timings, and configs[4] built on the device.
set -u
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-600}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
T=170 run python -m pytest tests/test_gpu_zz_scene_build.py tests/test_gpu_builder.py -x -q
T=60 run python tools/scene_build_bench.py 0 1000000
T=150 run python bench.py --scene soup --builder B200_PLOC --rays 33554432 --steps 3 --warmup 3 --cpu-seconds 0 2>gpurun_out/r02c19_bench_soup_ploc_resident.err | tee gpurun_out/r02c19_bench_soup_ploc_resident.json | cut -c1-200
