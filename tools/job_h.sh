#!/usr/bin/env bash
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-other 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('headline', d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
