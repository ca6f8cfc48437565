#!/usr/bin/env bash
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "assim or level_routing" 2>&1 | tail -2
tools/run_variants.sh f --skip-other 2>&1
one() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-other 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'])"; }
(cd variants/t_a05e2ca && one a05)
NCU="ncu --set full --clock-control none --import-source on -k regex:hru_sweep -s 2 -c 1"
python tools/prof_shapes.py blended_hourly_grid 1 50000 6 | tee gpurun_out/r2_prof_cfg4.txt
$NCU -f -o /tmp/r2_cfg4 python tools/prof_shapes.py blended_hourly_grid 1 50000 6 > gpurun_out/r2_ncu_cfg4.log 2>&1
python profiles/ncu_summary.py /tmp/r2_cfg4.ncu-rep > gpurun_out/r2_ncu_cfg4_summary.txt 2>&1
head -30 gpurun_out/r2_ncu_cfg4_summary.txt
