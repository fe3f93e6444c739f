set -x
python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_sparse.py -q -x 2>&1 | grep -v Warn | tail -4
VT_PROBE_UNWEIGHTED=1 VT_OGEMM_TIMING=1 python tools/syrk_probe.py 8000000 320 16000000 256 4000000 512 2000000 1024 | cut -c1-900
VT_OGEMM_TIMING=1 python tools/syrk_probe.py 2000000 1024 | cut -c1-900
