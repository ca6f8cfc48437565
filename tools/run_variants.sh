#!/usr/bin/env bash
# On the GPU box: run bench.py once per engine variant in variants/ and print kernel time + roofline fraction.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for so in variants/*.so; do
  n=$(basename $so .so)
  RVN_GPU_LIB=$PWD/$so timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err
  python - "$n" gpurun_out/var_$n.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); r=d["roofline"]
    print("%-12s sweep %.3f ms  step %.3f ms  frac %.3f  e2e %.3f ms" % (sys.argv[1], r["kernel_ms_per_launch"], d["ms_per_step"], r["frac"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
