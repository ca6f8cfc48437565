#!/usr/bin/env bash
# Build engine variants with different compile-time tuning macros into variants/<name>.so (kernel tuning experiments).
# usage: tools/build_variants.sh name1="-DA=1 -DB=2" name2="..."
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
mkdir -p "$ROOT/variants"
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -cudart static \
       -diag-suppress 550 -I"$ROOT/include" $flags "$ROOT/ravenhydroframework_b200/csrc/rvn_engine.cu" -o "$ROOT/variants/$name.so" &
done
wait
ls -la "$ROOT/variants"
