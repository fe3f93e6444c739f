set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --n-total 2000000 --steps 2 --warmup 1 --e2e-steps 1 > gpurun_out/bench_2gpu_2m.json 2> gpurun_out/bench_2gpu_2m.err
tail -5 gpurun_out/bench_2gpu_2m.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2gpu_2m.json').read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus','e2e','e2e_full','configs')}, indent=1)[:12000])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --cpu-sample 100000
