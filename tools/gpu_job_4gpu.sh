#!/bin/bash
# Four ranks on one box: the driver's scaling command for N = 4 (and the memory stations of the run on stderr).
set -x
free -g | head -2
VT_BENCH_MEMLOG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err
grep "\[mem\] rank 0" gpurun_out/bench_4gpu.err; tail -c 300 gpurun_out/bench_4gpu.err
