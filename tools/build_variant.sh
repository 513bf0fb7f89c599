#!/bin/bash
# Builds a differently compiled copy of the two libraries into luxcore_b200/lib_variants/<name>/ for A/B runs
# on the GPU box (LRB_LIB_DIR=<that directory> python ...).  Usage: tools/build_variant.sh <name> [-D...]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
out=$ROOT/luxcore_b200/lib_variants/$name
mkdir -p "$out"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
	-Xcompiler -fPIC,-fvisibility=hidden "$@" -shared -I"$ROOT/include" -I"$ROOT/luxcore_b200/csrc" \
	-o "$out/libluxrays_b200.so" "$ROOT/luxcore_b200/csrc/device.cu" "$ROOT/luxcore_b200/csrc/relayout.cpp"
g++ -O3 -std=c++17 -fPIC -shared -pthread -msse -msse2 -msse3 -mssse3 -ffp-contract=off -I"$ROOT/include" -I"$ROOT/luxcore_b200/host" \
	-o "$out/libluxrays_b200_host.so" "$ROOT"/luxcore_b200/host/*.cpp -L"$out" -lluxrays_b200 -Wl,-rpath,'$ORIGIN'
echo "built $out"
