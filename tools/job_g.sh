#!/usr/bin/env bash
python -m pytest tests/test_gpu_parity.py tests/test_dds.py tests/test_enkf.py -m gpu -q -x 2>&1 | tail -3
python tools/prof_shapes.py interp 128 500 4
cp ravenhydroframework_b200/librvn_gpu.so /tmp/keep.so; cp variants/interp_mb4.so ravenhydroframework_b200/librvn_gpu.so
python tools/prof_shapes.py interp 128 500 4
cp /tmp/keep.so ravenhydroframework_b200/librvn_gpu.so
python tools/bench_routing.py > gpurun_out/r2_routing3.json 2> gpurun_out/r2_routing3.err; cat gpurun_out/r2_routing3.json
