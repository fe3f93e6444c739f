set -x
mkdir -p gpurun_out
free -g | head -2; cat /sys/fs/cgroup/memory.max 2>/dev/null; cat /sys/fs/cgroup/memory/memory.limit_in_bytes 2>/dev/null
python -m pytest tests/test_gpu_solver.py -q -k "one_hyperparameter" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r02_final_1gpu.json 2> gpurun_out/bench_final.err; tail -3 gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_final_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'], d['roofline']['gemm_kernel_alone'])
print(json.dumps(d['kernels'])[:700])
print(json.dumps(d['e2e'])[:900]); print(json.dumps(d['e2e_full'])[:500])
for k,v in d['configs'].items():
    if isinstance(v, dict): print(k, json.dumps(v.get('step', v))[:300])
print(json.dumps(d['configs']['config4'])[:1200])
PY
