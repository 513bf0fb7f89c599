#!/bin/bash
# Round-2 opening GPU call (one box, one GPU): everything that was changed or prepared after the
# round-1 GPU budget ran out, measured in one go.  Usage:
#   gpurun --timeout 2400 -- 'bash tools/r02_first_call.sh > gpurun_out/r02_first_call.log 2>&1'
# Every step runs under its own timeout; outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
run() { echo; echo "=== $*"; timeout "${T:-600}" "$@"; echo "--- exit $?"; }

# 1. parity (the MBVH kernels changed: EnterInstance; the prefetching twin is new)
T=1500 run python -m pytest tests -m gpu -x -q

# 2. headline bench: new tree (default) vs the round-1 tree
T=600 run python bench.py --steps 10 --warmup 3
echo "--- round-1 builder (LRB_BVH_OPT=0)"
LRB_BVH_OPT=0 T=400 run python bench.py --steps 10 --warmup 3 --no-cpu-baseline

# 3. launch list + one full ncu capture of the headline kernel (never a bench value)
T=600 run ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
	python bench.py --steps 2 --warmup 3 --no-cpu-baseline
T=900 run ncu --set full --clock-control none --import-source on -k regex:TracePersistent -s 6 -c 1 -o gpurun_out/r02_kitchen \
	python bench.py --steps 2 --warmup 3 --no-cpu-baseline
T=200 run python tools/ncu_summary.py gpurun_out/r02_kitchen.ncu-rep

# 4. the soup after the axis-parallel fix: 16 M triangles (2 GB) then the full 50 M, sort / prefetch A/B
for sc in soup:16000000 soup; do
	for o in "sort_rays=0,prefetch=0" "sort_rays=1,prefetch=0" "sort_rays=0,prefetch=1" "sort_rays=1,prefetch=1"; do
		args=""; IFS=',' read -ra parts <<< "$o"; for p in "${parts[@]}"; do args="$args --opt $p"; done
		echo "--- $sc $o"
		T=900 run python bench.py --scene $sc --rays 33554432 --steps 3 --warmup 3 --no-cpu-baseline $args
	done
done
T=900 run ncu --set full --clock-control none --import-source on -k regex:TracePersistent -s 4 -c 1 -o gpurun_out/r02_soup16m \
	python bench.py --scene soup:16000000 --rays 33554432 --steps 1 --warmup 3 --no-cpu-baseline
T=200 run python tools/ncu_summary.py gpurun_out/r02_soup16m.ncu-rep

# 4b. two-level scenes: instance-entry vote (inst_bias 0 = enter at once, 8 = default)
for ib in 0 4 8 16; do
	echo "--- lightinstances inst_bias=$ib"
	T=600 run python tools/config_table.py --only lightinstances --opt inst_bias=$ib
done

# 5. the other configurations (MBVH: instance entry out of the pop loop) + one MBVH capture
T=1200 run python tools/config_table.py
cp -f gpurun_out/config_table.json gpurun_out/r02_config_table.json 2>/dev/null
