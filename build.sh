#!/usr/bin/env bash
# Build the C-ABI shared library for sm_100a, in-tree (travels to the GPU box with the snapshot).
# Translation units compile in parallel (objects under build/, git-ignored), then link.
set -euo pipefail
cd "$(dirname "$0")"
PKG="skill-chaining-with-graphs_b200"
OUT="$PKG/libscg_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-Wall -Xptxas -v)
mkdir -p build
: > build.log
pids=()
objs=()
for src in "$PKG"/csrc/*.cu; do
    obj="build/$(basename "${src%.cu}").o"
    objs+=("$obj")
    # rebuild only what changed (any header change rebuilds everything)
    if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ -n "$(find "$PKG/csrc" include -name '*.h' -newer "$obj" -o -name '*.cuh' -newer "$obj" | head -1)" ]; then
        ( "$NVCC" "${FLAGS[@]}" -c -o "$obj" "$src" > "$obj.log" 2>&1 ) &
        pids+=($!)
    fi
done
rc=0
for p in "${pids[@]:-}"; do [ -z "$p" ] || wait "$p" || rc=1; done
cat build/*.o.log >> build.log 2>/dev/null || true
if [ $rc -ne 0 ]; then grep -E "error" -A3 build.log | head -60; exit 1; fi
"$NVCC" -gencode arch=compute_100a,code=sm_100a --shared -o "$OUT" "${objs[@]}" >> build.log 2>&1 || { tail -30 build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "Wall" | head -20 || true
echo "built $OUT"
