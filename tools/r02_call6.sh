#!/bin/bash
# Round-2 GPU call 6: gate inside the triangle record + reference-ordered records (parity of the whole GPU suite), PopSpec off;
# bench lines + same-commit captures for kitchen / lightinstances / soup; resident-block variants on the soup.
set -u
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-600}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
T=1200 run python -m pytest tests -m gpu -q -x
T=200 run python tools/r02_measure.py kitchen --quick --tag _c6
T=200 run python tools/r02_measure.py mbvh --quick --tag _c6
T=400 run python bench.py --steps 10 --warmup 3 2>gpurun_out/r02c6_bench_kitchen.err | tee gpurun_out/r02c6_bench_kitchen.json | cut -c1-300
T=400 run python bench.py --scene lightinstances --accel MBVH --depth 1 --rays 4194304 --steps 10 --warmup 3 2>gpurun_out/r02c6_bench_mbvh.err | tee gpurun_out/r02c6_bench_mbvh.json | cut -c1-300
T=600 run ncu --set full --clock-control none --import-source on -k regex:TracePersistent -s 6 -c 1 -f -o gpurun_out/r02c6_kitchen \
	python bench.py --steps 2 --warmup 3 --no-cpu-baseline
T=600 run ncu --set full --clock-control none --import-source on -k regex:TracePersistent -s 8 -c 1 -f -o gpurun_out/r02c6_lightinstances \
	python bench.py --scene lightinstances --accel MBVH --depth 1 --rays 4194304 --steps 2 --warmup 3 --no-cpu-baseline
T=600 run python bench.py --scene soup --rays 33554432 --steps 3 --warmup 3 --cpu-seconds 8 2>gpurun_out/r02c6_bench_soup.err | tee gpurun_out/r02c6_bench_soup.json | cut -c1-300
T=600 run ncu --set full --clock-control none --import-source on -k regex:TracePersistent -s 3 -c 1 -f -o gpurun_out/r02c6_soup \
	python tools/r02_measure.py soup --quick --no-parity --tag _c6_ncu
T=400 run python tools/r02_measure.py soup --quick --no-parity --opt blocks_per_sm=8 --tag _c6_b8
LRB_LIB_DIR=$PWD/luxcore_b200/lib_variants/b12 T=400 run python tools/r02_measure.py soup --quick --no-parity --tag _c6_b12
for f in kitchen lightinstances soup; do python tools/ncu_summary.py gpurun_out/r02c6_$f.ncu-rep > gpurun_out/r02c6_${f}_ncu_summary.txt 2>&1; done
ls -la gpurun_out | tail -12
