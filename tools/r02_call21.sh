#!/bin/bash
# Round-2 GPU call 21: tapered chunk schedule of the host pipelines (host_chunks.h) -- bench line with and without it
# (e2e now verified byte for byte against a device-resident trace), then the tests of the host pipelines.
set -u
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-600}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
T=40 run python bench.py 2>gpurun_out/r02c21_bench.err | tee gpurun_out/r02c21_bench.json | cut -c1-200
T=40 run python bench.py --opt host_taper=0 --no-cpu-baseline 2>gpurun_out/r02c21_bench_notaper.err | tee gpurun_out/r02c21_bench_notaper.json | cut -c1-200
T=75 run python -m pytest -x -q tests/test_gpu_pipeline.py tests/test_gpu_zz_scene_build.py::test_resident_scene_of_a_flattened_instance_scene tests/test_gpu_host_layer.py
