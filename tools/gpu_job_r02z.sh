set -x
mkdir -p gpurun_out
python tools/sanitize_run.py 2>&1 | tail -30
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_run.py > gpurun_out/sanitizer_new_kernels.log 2>&1; echo rc=$?
grep -c "Invalid" gpurun_out/sanitizer_new_kernels.log; grep "ERROR SUMMARY\|sanitize run done" gpurun_out/sanitizer_new_kernels.log; grep -m3 -A12 "Invalid" gpurun_out/sanitizer_new_kernels.log | cut -c1-200
python -m pytest tests/test_gpu_ozaki.py -q 2>&1 | tail -2
python tools/cmax_probe.py
