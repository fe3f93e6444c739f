# Round-1 closing job (run under gpurun): GPU test-suite, smoke, the full bench line (N = 10M) and the CPU reference arm.
set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/gpu_tests_final.log 2>&1
tail -4 gpurun_out/gpu_tests_final.log
python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -3 gpurun_out/bench_full.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2>> gpurun_out/bench_full.err
du -sh gpurun_out
