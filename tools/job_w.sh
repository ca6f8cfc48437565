#!/usr/bin/env bash
# launch list of the headline bench command (setup = 512 per-member transposes first, then the step kernels)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --skip-other > gpurun_out/r2_ncu_launches2.log 2>&1
python profiles/launch_summary.py gpurun_out/r2_launches2.csv transpose_in_kernel
