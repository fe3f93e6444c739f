# Round 2, final state: GPU test suite, smoke(), chol / block probes, default bench.
set -x
mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python tools/chol_probe.py > gpurun_out/chol_probe_r02.jsonl 2>&1; cat gpurun_out/chol_probe_r02.jsonl
python tools/block_probe.py 1000000 > gpurun_out/block_probe_r02.jsonl 2>&1; cat gpurun_out/block_probe_r02.jsonl
python bench.py > gpurun_out/bench_r02_final_1gpu.json 2> gpurun_out/bench_final.err; tail -3 gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_final_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'], d['roofline']['gemm_kernel_alone'])
print(json.dumps(d['kernels'])[:700])
print(json.dumps(d['e2e'])[:800]); print(json.dumps(d['e2e_full'])[:500])
for k,v in d['configs'].items():
    if isinstance(v, dict): print(k, json.dumps(v.get('step', v))[:300])
print(json.dumps(d['configs']['config4'])[:1200])
PY
