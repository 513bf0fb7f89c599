#!/bin/bash
# Round-2 GPU call 2: parity of everything (new any-hit / pass-through / compaction rows included), A/B of
# differently compiled kernels (resident blocks per SM), masked-batch and shadow-ray measurements.
set -u
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-600}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
T=900 run python -m pytest tests -m gpu -x -q
for v in b9 b10; do
	echo "--- variant $v"
	LRB_LIB_DIR=$PWD/luxcore_b200/lib_variants/$v T=300 run python tools/r02_measure.py kitchen --quick --tag _$v --opt carveout=100
done
T=300 run python tools/r02_measure.py kitchen --quick --tag _base
T=300 run python tools/r02_measure.py kitchen --quick --tag _base_c100 --opt carveout=100
for v in m7 m8; do
	echo "--- variant $v"
	LRB_LIB_DIR=$PWD/luxcore_b200/lib_variants/$v T=300 run python tools/r02_measure.py mbvh --quick --tag _$v
done
T=300 run python tools/r02_measure.py mbvh --quick --tag _base
T=600 run python tools/r02_measure.py masked
ls gpurun_out
