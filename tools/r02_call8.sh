#!/bin/bash
# Round-2 GPU call 8 (N GPUs): gather variants (signalled in-kernel pushes / whole-slice push after a plain kernel, both
# pipelined), NCCL gather, strong scaling, timeline; N >= 8 adds configs[4] (50 M-triangle soup, 32 Mi rays per GPU).
set -u
N=${1:-2}
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-300}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv,noheader | head -8
b() { local tag=$1; shift; T=300 run $TR bench.py --gpus $N --steps 10 --warmup 3 "$@" 2>gpurun_out/r02c8_n${N}_$tag.err | tee gpurun_out/r02c8_n${N}_$tag.json | cut -c1-160; }
if [ "$N" -ge 4 ]; then
	b push1 --chunks 1
	b signalled --chunks 0 --no-cpu-baseline
	b nccl --gather nccl --no-cpu-baseline
	b strong_push1 --scaling strong --chunks 1 --no-cpu-baseline
else
	b signalled --chunks 0 --timeline
	b push1 --chunks 1 --no-cpu-baseline
	b push4 --chunks 4 --no-cpu-baseline
	b nccl --gather nccl --no-cpu-baseline
	b strong --scaling strong --chunks 0 --no-cpu-baseline
	b strong_push1 --scaling strong --chunks 1 --no-cpu-baseline
fi
if [ "$N" -ge 8 ]; then
	T=900 run $TR bench.py --gpus $N --scene soup --rays 33554432 --steps 3 --warmup 3 --no-cpu-baseline --chunks 1 2>gpurun_out/r02c8_n${N}_soup.err | tee gpurun_out/r02c8_n${N}_soup.json | cut -c1-160
fi
tail -3 gpurun_out/r02c8_n${N}_*.err
