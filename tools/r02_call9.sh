#!/bin/bash
# Round-2 GPU call 9: A/B of the both-phases-per-iteration schedule (LRB_BOTH_PHASES = 1 / 4 / 8) against the voted phases.
set -u
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-600}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
T=200 run python tools/r02_measure.py kitchen --quick --tag _c9_vote
T=200 run python tools/r02_measure.py mbvh --quick --tag _c9_vote
for v in both1 both4 both8; do
	LRB_LIB_DIR=$PWD/luxcore_b200/lib_variants/$v T=200 run python tools/r02_measure.py kitchen --quick --tag _c9_$v
	LRB_LIB_DIR=$PWD/luxcore_b200/lib_variants/$v T=200 run python tools/r02_measure.py mbvh --quick --tag _c9_$v
done
