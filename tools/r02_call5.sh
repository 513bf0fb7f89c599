#!/bin/bash
# Round-2 GPU call 5: speculative pop at the end of the phases (LRB_POPSPEC 0 / 1 / 2), bottom-entry stack, 64-bit stack
# entries: parity of the default build, A/B of the variants, bench lines + a source-level capture of the new kernel.
set -u
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-600}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
T=200 run python tools/r02_measure.py kitchen --quick --tag _c5_ps1
for v in ps0 ps2 ps1b8; do
	LRB_LIB_DIR=$PWD/luxcore_b200/lib_variants/$v T=200 run python tools/r02_measure.py kitchen --quick --tag _c5_$v
done
T=200 run python tools/r02_measure.py mbvh --quick --tag _c5_ps1
for v in ps0 ps2; do
	LRB_LIB_DIR=$PWD/luxcore_b200/lib_variants/$v T=200 run python tools/r02_measure.py mbvh --quick --tag _c5_$v
done
T=1200 run python -m pytest tests -m gpu -q -x
T=400 run python bench.py --steps 10 --warmup 3 2>gpurun_out/r02c5_bench_kitchen.err | tee gpurun_out/r02c5_bench_kitchen.json | cut -c1-300
T=400 run python bench.py --scene lightinstances --accel MBVH --depth 1 --rays 4194304 --steps 10 --warmup 3 2>gpurun_out/r02c5_bench_mbvh.err | tee gpurun_out/r02c5_bench_mbvh.json | cut -c1-300
T=600 run ncu --set full --clock-control none --import-source on -k regex:TracePersistent -s 6 -c 1 -f -o gpurun_out/r02c5_kitchen \
	python bench.py --steps 2 --warmup 3 --no-cpu-baseline
T=600 run ncu --set full --clock-control none --import-source on -k regex:TracePersistent -s 8 -c 1 -f -o gpurun_out/r02c5_lightinstances \
	python bench.py --scene lightinstances --accel MBVH --depth 1 --rays 4194304 --steps 2 --warmup 3 --no-cpu-baseline
ls -la gpurun_out | tail -12
