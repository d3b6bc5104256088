#!/bin/bash
# developer tool: scripts/build_variant.sh <name> <extra nvcc flags...> -> scripts/libqot_b200_<name>.so
# (same sources, extra -D switches; select with QOT_B200_LIB=scripts/libqot_b200_<name>.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
objs=""
for f in gnn_qot_estimation_b200/build/*.o; do
  case "$f" in *lightpath_stream.o) ;; *) objs="$objs $f";; esac
done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr \
  -Xcompiler -fPIC "$@" -c gnn_qot_estimation_b200/csrc/lightpath_stream.cu -o /tmp/lightpath_stream_$name.o -Xptxas -v 2>&1 | grep -A2 "lp_stream_kernel" | grep "registers\|spill" || true
nvcc -shared -o scripts/libqot_b200_$name.so $objs /tmp/lightpath_stream_$name.o -gencode arch=compute_100a,code=sm_100a -cudart static
echo built scripts/libqot_b200_$name.so
