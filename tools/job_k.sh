#!/usr/bin/env bash
one() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-other 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'])"; }
one current_a; (cd variants/t_a05e2ca && one a05); (cd variants/t_pad && one pad); one current_b
