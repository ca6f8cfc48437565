#!/usr/bin/env bash
python -m pytest tests/test_gpu_parity.py tests/test_reference_inputs.py -m gpu -q -x 2>&1 | tail -3
tools/run_variants.sh e --skip-other 2>&1
one() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-other 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'])"; }
(cd variants/t_a05e2ca && one a05)
python tools/bench_routing.py > gpurun_out/r2_routing5.json 2> gpurun_out/r2_routing5.err; cat gpurun_out/r2_routing5.json
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --only 4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); c=d['other_configs']['4']; print('cfg4', c['value'], c['ms_per_step'], c['launches_per_step'], c['roofline']['kernel_ms_per_launch'])"
