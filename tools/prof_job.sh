#!/usr/bin/env bash
# Development helper (GPU box): ncu captures of the sweep kernels of every named shape; the reports are summarised on the box
# (profiles/ncu_summary.py) because gpurun only brings back 64 MiB.
NCU="ncu --set full --clock-control none --import-source on -k regex:hru_sweep -s 1 -c 1"
run() {  # name, args...
  name="$1"; shift
  python tools/prof_shapes.py "$@" | tee "gpurun_out/r2_prof_${name}.txt"
  $NCU -f -o "/tmp/r2_${name}" python tools/prof_shapes.py "$@" > "gpurun_out/r2_ncu_${name}.log" 2>&1
  python profiles/ncu_summary.py "/tmp/r2_${name}.ncu-rep" > "gpurun_out/r2_ncu_${name}_summary.txt" 2>&1
}
run hbvec hbv 128 5000 16
run blended blended 1 20000 8
run gr4j gr4j 512 500 8
run interp interp 128 500 4
run hmets hmets 512 500 4
ls -la gpurun_out | head -40
# config 4 at full size: hourly Blended on the 200 x 200 grid, 1 M HRUs
NCU="ncu --set full --clock-control none --import-source on -k regex:hru_sweep -s 2 -c 1"
python tools/prof_shapes.py blended_hourly_grid 1 50000 6 | tee gpurun_out/r2_prof_cfg4.txt
$NCU -f -o /tmp/r2_cfg4 python tools/prof_shapes.py blended_hourly_grid 1 50000 6 > gpurun_out/r2_ncu_cfg4.log 2>&1
python profiles/ncu_summary.py /tmp/r2_cfg4.ncu-rep > gpurun_out/r2_ncu_cfg4_summary.txt 2>&1
