set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-e2e-full > gpurun_out/bench_r03g.json 2> gpurun_out/bench_r03g.err
tail -c 600 gpurun_out/bench_r03g.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r03g.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')})
print(d['kernels'])
print(d['roofline']['frac'], d['roofline']['gemm_kernel_alone']['frac'])
print(d['e2e']['value'])
for c in ('config3', 'config4', 'config5'):
    print(c, d['configs'][c]['step'])
PY
