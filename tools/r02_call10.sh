#!/bin/bash
# Round-2 GPU call 10: L2 persistence window (l2_persist) alone and under emulated gather ingest; the soup parity test;
# GPU test suite on the slab layout.
set -u
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-600}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
T=300 run python tools/r02_measure.py ingest --tag _c10
T=200 run python tools/r02_measure.py kitchen --quick --opt l2_persist=0 --tag _c10_p0
T=200 run python tools/r02_measure.py kitchen --quick --opt l2_persist=2 --tag _c10_p2
T=900 run python -m pytest tests/test_gpu_zz_soup.py -q -x
T=1200 run python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_zz_soup.py
