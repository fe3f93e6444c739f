#!/bin/bash
# End-of-round validation on one B200: GPU tests, smoke, the default bench line and the reference arm.
set -x
python tools/chol_probe.py > gpurun_out/chol_probe.jsonl; cut -c1-200 gpurun_out/chol_probe.jsonl
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 300 gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_final.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')})
print({k: (v['ms'] if isinstance(v, dict) and 'ms' in v else v) for k, v in d['kernels'].items()})
print('roofline', d['roofline']['frac'], d['roofline']['gemm_kernel_alone']['frac'])
print('e2e', d['e2e']['value'], d.get('e2e_full'))
for c in ('config3', 'config4', 'config5'):
    print(c, d['configs'][c]['step'])
print({k: v['ms'] for k, v in d['configs']['config3'].items() if isinstance(v, dict) and 'ms' in v})
print(d['configs']['config4']['potrf'], d['configs']['config4']['potrs'])
r = json.load(open('gpurun_out/bench_reference.json'))
print('reference', r['value'], r['cpu_baseline'])
PY
