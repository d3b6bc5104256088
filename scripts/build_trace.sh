#!/bin/bash
# instrumented build of the same sources (developer tool, not shipped): -DQOT_LP_TRACE
set -e
cd "$(dirname "$0")/.."

nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr \
  -Xcompiler -fPIC -DQOT_LP_TRACE -shared -cudart static \
  gnn_qot_estimation_b200/csrc/*.cu -o scripts/libqot_b200_trace.so
