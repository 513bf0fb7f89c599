#!/bin/bash
# Round-2 GPU call 4 (N GPUs, default 2): pipelined / un-pipelined gather, timeline, strong scaling, film merge.
set -u
N=${1:-2}
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-300}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv,noheader | head -8
T=300 run $TR bench.py --gpus $N --steps 10 --warmup 3 --timeline 2>gpurun_out/r02c4_n${N}_pipe.err | tee gpurun_out/r02c4_n${N}_pipe.json | cut -c1-200
T=300 run $TR bench.py --gpus $N --steps 10 --warmup 3 --no-pipeline --no-cpu-baseline 2>gpurun_out/r02c4_n${N}_nopipe.err | tee gpurun_out/r02c4_n${N}_nopipe.json | cut -c1-200
T=300 run $TR bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-cpu-baseline 2>gpurun_out/r02c4_n${N}_strong.err | tee gpurun_out/r02c4_n${N}_strong.json | cut -c1-200
T=300 run $TR bench.py --gpus $N --steps 10 --warmup 3 --gather nccl --no-cpu-baseline 2>gpurun_out/r02c4_n${N}_nccl.err | tee gpurun_out/r02c4_n${N}_nccl.json | cut -c1-200
T=200 run $TR tools/film_merge_check.py
if [ "$N" = "2" ]; then
	T=900 run python -m pytest tests -m gpu -q -k "advance or film or in_plane or reference_library or gather"
fi
tail -5 gpurun_out/r02c4_n${N}_*.err
