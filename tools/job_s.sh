#!/usr/bin/env bash
# the driver's 8-GPU command on the final tree
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
echo "rc=$?"; head -c 600 gpurun_out/r2_bench_8gpu.json; tail -3 gpurun_out/r2_bench_8gpu.err
