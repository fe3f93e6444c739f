set -x
python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_ij.py tests/test_gpu_taylor.py -q 2>&1 | tail -8
python bench.py --n-total 2000000 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-configs --no-tf32 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], json.dumps(d['kernels'])[:600])"
