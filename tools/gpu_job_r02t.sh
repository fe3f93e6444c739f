set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --no-tf32 > gpurun_out/bench_r02_8gpu.json 2> gpurun_out/bench_8gpu.err
tail -5 gpurun_out/bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_8gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['n_gpus'], d['clocks'])
print(json.dumps(d['e2e'])[:900]); print(json.dumps(d['e2e_full'])[:500])
for k,v in d['configs'].items():
    if isinstance(v, dict): print(k, json.dumps(v.get('step', v))[:300])
PY
