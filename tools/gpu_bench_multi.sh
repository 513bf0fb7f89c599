#!/bin/bash
# usage: tools/gpu_bench_multi.sh N "gather args" ["gather args" ...]   -- every run under a hard timeout
N=$1; shift
for g in "$@"; do
  echo "== N=$N $g"
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --gather $g 2>gpurun_out/multi.err | python -c "
import json,sys
t=[l for l in sys.stdin.read().splitlines() if l.startswith('{')]
if not t: print('NO OUTPUT'); sys.exit(0)
d=json.loads(t[-1]); print(d['value'], 'Mrays/s  ms/step', d['ms_per_step'], d['gather'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'])"
  grep -i -E "error|Traceback|watchdog" gpurun_out/multi.err | head -3
done
