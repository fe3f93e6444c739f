# Round 2: the GPU test suite, smoke(), and the default bench + reference arm, as the driver runs them.
set -x
mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference > gpurun_out/bench_r02_final_reference_arm.json 2> gpurun_out/bench_ref.err; tail -2 gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_r02_final_1gpu.json 2> gpurun_out/bench_final.err; tail -3 gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_final_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'])
print(json.dumps(d['kernels'])[:700])
print(json.dumps(d['e2e'])[:800]); print(json.dumps(d['e2e_full'])[:500])
for k,v in d['configs'].items():
    if isinstance(v, dict): print(k, json.dumps(v.get('step', v))[:300])
r=json.loads(open('gpurun_out/bench_r02_final_reference_arm.json').read().strip().splitlines()[-1])
print('reference', r['value'], r['cpu_baseline']['cores'], r['config'])
PY
