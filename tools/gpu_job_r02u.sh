set -x
mkdir -p gpurun_out
echo ring4; timeout 150 python tools/ogemm_probe.py time_parts 2>&1 | grep '"S": 7'
echo ring6; VT_LIB_PATH=$PWD/vittles_b200/lib/libvittles_b200_ring6.so timeout 150 python tools/ogemm_probe.py time_parts 2>&1 | grep '"S": 7'
echo ring4 apply; timeout 150 python tools/ogemm_probe.py time_apply; timeout 150 python tools/ogemm_probe.py time_syrk
echo ring6 apply; VT_LIB_PATH=$PWD/vittles_b200/lib/libvittles_b200_ring6.so timeout 150 python tools/ogemm_probe.py time_apply; VT_LIB_PATH=$PWD/vittles_b200/lib/libvittles_b200_ring6.so timeout 150 python tools/ogemm_probe.py time_syrk
python bench.py --n-total 1000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-tf32 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], json.dumps(d['roofline'])[:1800])"
