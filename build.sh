#!/usr/bin/env bash
# Build the C-ABI shared library for sm_100a, in-tree (travels to the GPU box with the snapshot).
set -euo pipefail
cd "$(dirname "$0")"
PKG="skill-chaining-with-graphs_b200"
OUT="$PKG/libscg_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -Xcompiler -fPIC,-ffp-contract=off,-Wall -Xptxas -v --shared \
    -o "$OUT" "$PKG"/csrc/scg_api.cu "$PKG"/csrc/scg_step.cu "$PKG"/csrc/scg_q.cu \
    "$PKG"/csrc/scg_sarsa.cu "$PKG"/csrc/scg_agent.cu "$PKG"/csrc/scg_xchg.cu 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "Wall" | head -20 || true
echo "built $OUT"
