#!/bin/bash
# Round-2 GPU call 14 (2 GPUs): PLOC builder (tests, times, kitchen traversal), flag-based completion signal of the gather.
set -u
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-300}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
T=600 run python -m pytest tests/test_gpu_builder.py -q -x
T=300 run python tools/builder_bench.py
T=300 run python bench.py --builder B200_PLOC --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02c14_bench_kitchen_ploc.err | tee gpurun_out/r02c14_bench_kitchen_ploc.json | cut -c1-200
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
T=300 run $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02c14_n2_flag.err | tee gpurun_out/r02c14_n2_flag.json | cut -c1-200
T=300 run $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --signal nccl 2>gpurun_out/r02c14_n2_nccl_signal.err | tee gpurun_out/r02c14_n2_nccl_signal.json | cut -c1-200
T=300 run $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --scaling strong 2>gpurun_out/r02c14_n2_strong_flag.err | tee gpurun_out/r02c14_n2_strong_flag.json | cut -c1-200
tail -n 3 gpurun_out/r02c14_n2_*.err
