#!/bin/bash
# Round-2 GPU call 3: parity on the new default build (10 / 8 resident blocks, re-written pop loop), bench lines + captures.
set -u
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-600}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
T=1200 run python -m pytest tests -m gpu -q
T=400 run python bench.py --steps 10 --warmup 3 2>gpurun_out/r02c3_bench_kitchen.err | tee gpurun_out/r02c3_bench_kitchen.json | cut -c1-300
T=400 run python bench.py --scene lightinstances --accel MBVH --depth 1 --rays 4194304 --steps 10 --warmup 3 2>gpurun_out/r02c3_bench_mbvh.err | tee gpurun_out/r02c3_bench_mbvh.json | cut -c1-300
T=600 run ncu --set full --clock-control none --import-source on -k regex:TracePersistent -s 6 -c 1 -f -o gpurun_out/r02c3_kitchen \
	python bench.py --steps 2 --warmup 3 --no-cpu-baseline
T=600 run ncu --set full --clock-control none --import-source on -k regex:TracePersistent -s 8 -c 1 -f -o gpurun_out/r02c3_lightinstances \
	python bench.py --scene lightinstances --accel MBVH --depth 1 --rays 4194304 --steps 2 --warmup 3 --no-cpu-baseline
T=300 run python tools/r02_measure.py masked --tag _auto
ls -la gpurun_out | tail -12
