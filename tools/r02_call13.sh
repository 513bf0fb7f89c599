#!/bin/bash
# Round-2 GPU call 13 (N GPUs): final scaling lines with every rank's trace time inside the gather pipeline.
set -u
N=${1:-8}
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-300}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
b() { local tag=$1; shift; T=300 run $TR bench.py --gpus $N --steps 10 --warmup 3 "$@" 2>gpurun_out/r02c13_n${N}_$tag.err | tee gpurun_out/r02c13_n${N}_$tag.json | cut -c1-160; }
b weak --no-cpu-baseline
b strong --scaling strong --no-cpu-baseline
if [ "$N" -ge 8 ]; then
	# four GPUs of the same box, for the scaling table
	T=300 run python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02c13_n4_weak.err | tee gpurun_out/r02c13_n4_weak.json | cut -c1-160
	T=300 run python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 --scaling strong --no-cpu-baseline 2>gpurun_out/r02c13_n4_strong.err | tee gpurun_out/r02c13_n4_strong.json | cut -c1-160
fi
