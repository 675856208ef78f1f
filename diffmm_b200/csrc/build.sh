#!/usr/bin/env bash
# Builds libdiffmm_b200.so in-tree for sm_100a.  nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libdiffmm_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr
       -Xptxas -v -I"${HERE}/../../include")
mkdir -p "${HERE}/obj"
pids=()
for f in capi gemm_tcgen05 gemm_simt pack topk adj spmm prop optim loss eval train; do
  src="${HERE}/${f}.cu"; obj="${HERE}/obj/${f}.o"
  if [[ ! -f "$obj" || "$src" -nt "$obj" || "${HERE}/common.cuh" -nt "$obj" || "${HERE}/../../include/diffmm_b200.h" -nt "$obj" ]]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$src" -o "$obj" > "${HERE}/obj/${f}.log" 2>&1 || { cat "${HERE}/obj/${f}.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -shared -o "$OUT" "${HERE}"/obj/*.o -lcudart
echo "built $OUT"
