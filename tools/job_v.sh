#!/usr/bin/env bash
# dress rehearsal of what the driver runs at round end: GPU tests, smoke, both bench arms
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 200 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc=$?"; head -c 400 gpurun_out/r2_final_bench.json
