#!/usr/bin/env bash
python -m pytest tests/test_gpu_parity.py tests/test_reference_inputs.py tests/test_checkpoint.py -m gpu -q -x -k "res_ or diffusive or LOTW or Nith or reservoirs or hydrologic or muskingum" 2>&1 | tail -5
python tools/bench_routing.py > gpurun_out/r2_routing4.json 2> gpurun_out/r2_routing4.err; cat gpurun_out/r2_routing4.json
