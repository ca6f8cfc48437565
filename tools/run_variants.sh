#!/usr/bin/env bash
# Development helper (GPU box): bench every variants/*.so (kernel-tuning builds of the engine) with the default workload.
# usage: tools/run_variants.sh <tag> [bench args]
tag="$1"; shift
cp ravenhydroframework_b200/librvn_gpu.so /tmp/keep.so
for v in variants/*.so; do
  n=$(basename "$v" .so)
  cp "$v" ravenhydroframework_b200/librvn_gpu.so
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" > "gpurun_out/var_${tag}_${n}.json" 2> "gpurun_out/var_${tag}_${n}.err"
  python - "$n" "gpurun_out/var_${tag}_${n}.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    r = d["roofline"]
    print("%-24s ms/step %.4f  sweep %.4f  frac %.4f  e2e %.4f ms" % (sys.argv[1], d["ms_per_step"], r["kernel_ms_per_launch"], r["frac"], d["e2e"]["ms_per_step"]))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
cp /tmp/keep.so ravenhydroframework_b200/librvn_gpu.so
