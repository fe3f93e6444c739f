#!/bin/bash
# Two ranks on one box: the sharded paths of the bench (NCCL) and the launch list of one factorisation + solve.
set -x
VT_BENCH_MEMLOG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
grep "\[mem\]" gpurun_out/bench_2gpu.err; free -g; tail -c 300 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_2gpu.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'gpu_launches')})
print('e2e', d['e2e'].get('value'), 'e2e_full', (d.get('e2e_full') or {}).get('value'))
for c in ('config3', 'config4', 'config5'):
    print(c, d['configs'][c]['step'])
PY
