python -m pytest tests/test_gpu_solver.py tests/test_gpu_lrcov.py tests/test_gpu_ij.py -q -x 2>&1 | grep -v Warning | tail -45
