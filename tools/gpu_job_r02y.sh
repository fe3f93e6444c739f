# memcheck of the round's new kernels on small test cases
set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_ozaki.py -q -x -k "digits_and_products or hands_the_hessian or ij_apply_on_the_int8 or extreme" > gpurun_out/sanitizer_ozaki.log 2>&1; echo rc=$?; tail -5 gpurun_out/sanitizer_ozaki.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_solver.py tests/test_gpu_sparse.py -q -x -k "cg_columns or preconditioner or many_right or block_kernels or wider_than" > gpurun_out/sanitizer_solver.log 2>&1; echo rc=$?; tail -5 gpurun_out/sanitizer_solver.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/chol_probe.py > gpurun_out/sanitizer_chol.log 2>&1; echo rc=$?; tail -4 gpurun_out/sanitizer_chol.log | cut -c1-300
grep -c "Invalid\|ERROR SUMMARY" gpurun_out/sanitizer_*.log; grep "ERROR SUMMARY" gpurun_out/sanitizer_*.log
