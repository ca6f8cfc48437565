#!/usr/bin/env bash
# headline bench, ncu full capture + launch list of the same command, then the complete bench line (every named shape, CPU baseline)
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-other > gpurun_out/r2_bench_head.json 2> gpurun_out/r2_bench_head.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:hru_sweep_static -s 6 -c 1 -f -o /tmp/r2_hmets python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-other > gpurun_out/r2_ncu_hmets2.log 2>&1
python profiles/ncu_summary.py /tmp/r2_hmets.ncu-rep > gpurun_out/r2_ncu_hmets2_summary.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --skip-other > gpurun_out/r2_ncu_launches.log 2>&1
python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err
head -c 600 gpurun_out/r2_bench_full.json; echo; head -12 gpurun_out/r2_ncu_hmets2_summary.txt
