#!/bin/bash
# usage (on the GPU box): tools/gpu_bench_variants.sh "opt1" "opt2,opt3" ...   (comma separates several --opt)
for o in "$@"; do
  args=""
  IFS=',' read -ra parts <<< "$o"
  for p in "${parts[@]}"; do [ -n "$p" ] && [ "$p" != "default" ] && args="$args --opt $p"; done
  echo "== $o"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline $args 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print(d['value'], 'Mrays/s kernel_ms', r['kernel_ms'], 'nodes/ray', r['impl_wide_nodes_per_ray'], 'tris/ray', r['impl_triangles_per_ray'], 'e2e', d['e2e']['value'], 'parity', d.get('parity_check'))"
done
