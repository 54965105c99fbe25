#!/usr/bin/env bash
# Builds the sm_100a CUDA library in-tree (same command __graft_entry__.build() runs).
set -euo pipefail
cd "$(dirname "$0")/differentiable_tube_mpc_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
     -shared -Xcompiler -fPIC -o ../libdtmpc_b200.so dtmpc.cu "$@"
