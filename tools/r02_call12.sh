#!/bin/bash
# Round-2 GPU call 12: what the prefetching kernel should fetch ahead on the 50 M-triangle soup (prefetch_mode sweep),
# upload time with the multi-threaded re-layout; memcheck of the builder and of a small trace.
set -u
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-600}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
T=600 run python tools/r02_measure.py soup --no-parity --prefetch-modes --tag _c12_prefetch
T=300 run python tools/r02_measure.py kitchen --quick --opt prefetch=1 --opt prefetch_mode=2 --tag _c12_pf2
T=300 run python tools/r02_measure.py kitchen --quick --opt prefetch=1 --opt prefetch_mode=4 --tag _c12_pf4
T=600 run compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_builder.py -q -x -k "array_rules and not 200000"
T=600 run compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_bvh.py -q -x -k "cornell"
