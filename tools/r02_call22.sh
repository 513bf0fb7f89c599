#!/bin/bash
# Round-2 GPU call 22: the host pipelines' tests with the tapered chunk schedule + the flattened-instance resident scene.
set -u
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-600}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
T=85 run python -m pytest -q tests/test_gpu_pipeline.py tests/test_gpu_zz_scene_build.py::test_resident_scene_of_a_flattened_instance_scene tests/test_gpu_host_layer.py
