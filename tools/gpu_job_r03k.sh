set -x
python -m pytest tests/test_gpu_ozaki.py -q -x 2>&1 | tail -2
VT_OGEMM_TIMING=1 python tools/ogemm_probe.py timing 2>&1 | tail -2
VT_OGEMM_TIMING=1 python tools/syrk_probe.py 2000000 1024 | cut -c1-200
python tools/chol_probe.py | cut -c1-330
