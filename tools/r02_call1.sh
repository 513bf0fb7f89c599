#!/bin/bash
# Round-2 GPU call 1 (one GPU): state of HEAD on every configuration + ncu captures.
#   gpurun --timeout 2100 -- 'bash tools/r02_call1.sh > gpurun_out/r02_call1.log 2>&1'
set -u
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-600}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv
nproc; free -g | head -2

# 1. headline bench line of HEAD + launch list + full capture of the timed kernel
T=400 run python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_head.json
tail -c 3000 gpurun_out/r02_bench_head.json
T=300 run ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_kitchen.csv \
	python bench.py --steps 2 --warmup 3 --no-cpu-baseline
T=600 run ncu --set full --clock-control none --import-source on -k regex:TracePersistent -s 6 -c 1 -f -o gpurun_out/r02_kitchen \
	python bench.py --steps 2 --warmup 3 --no-cpu-baseline
T=120 run python tools/ncu_summary.py gpurun_out/r02_kitchen.ncu-rep > gpurun_out/r02_kitchen_ncu_summary.txt

# 2. option sweeps on the headline workload (+ classroom)
T=600 run python tools/r02_measure.py kitchen

# 3. configs[4]: the 50 M-triangle soup -- sweeps, bench line, full capture
T=900 run python tools/r02_measure.py soup
T=900 run python bench.py --scene soup --rays 33554432 --steps 3 --warmup 3 --cpu-seconds 8 > gpurun_out/r02_bench_soup.json
tail -c 3000 gpurun_out/r02_bench_soup.json
T=900 run ncu --set full --clock-control none --import-source on -k regex:TracePersistent -s 4 -c 1 -f -o gpurun_out/r02_soup \
	python bench.py --scene soup --rays 33554432 --steps 1 --warmup 3 --no-cpu-baseline
T=120 run python tools/ncu_summary.py gpurun_out/r02_soup.ncu-rep > gpurun_out/r02_soup_ncu_summary.txt

# 4. two-level scenes: sweeps + one full capture of a lightinstances bounce-1 launch
T=900 run python tools/r02_measure.py mbvh
T=600 run ncu --set full --clock-control none --import-source on -k regex:TracePersistent -s 12 -c 1 -f -o gpurun_out/r02_lightinstances \
	python tools/config_table.py --only lightinstances
T=120 run python tools/ncu_summary.py gpurun_out/r02_lightinstances.ncu-rep > gpurun_out/r02_lightinstances_ncu_summary.txt
ls -la gpurun_out
