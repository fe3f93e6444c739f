#!/bin/bash
# Eight ranks on one box: the driver's scaling command for N = 8 (and the memory stations of the run on stderr).
set -x
free -g | head -2
VT_BENCH_MEMLOG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
grep "\[mem\] rank 0" gpurun_out/bench_8gpu.err; tail -c 300 gpurun_out/bench_8gpu.err
