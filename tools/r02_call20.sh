#!/bin/bash
# Round-2 GPU call 20: the whole GPU suite and smoke() after the re-layout's per-node code moved into relayout_shared.h
# (every upload goes through it) and lrb_bvh_build_scene joined the library.
set -u
mkdir -p gpurun_out
run() { echo; echo "=== $*"; local t0=$SECONDS; timeout "${T:-600}" "$@"; echo "--- exit $? after $((SECONDS-t0)) s"; }
T=300 run python -m pytest tests -m gpu -q -x
T=60 run python __graft_entry__.py smoke
