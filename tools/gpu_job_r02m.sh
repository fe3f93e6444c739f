set -x
mkdir -p gpurun_out
python tools/chol_probe.py > gpurun_out/chol_probe_r02m.jsonl 2>&1; cat gpurun_out/chol_probe_r02m.jsonl
python bench.py --n-total 1000000 --steps 2 --warmup 1 --e2e-steps 1 --no-tf32 > gpurun_out/bench_1m_r02m.json 2> gpurun_out/bench_1m_r02m.err
tail -5 gpurun_out/bench_1m_r02m.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_1m_r02m.json').read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','e2e_full','configs')}, indent=1)[:9000])
PY
