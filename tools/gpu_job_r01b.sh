# Round-1 measurement job (run under gpurun): full bench line, ncu launch list of the bench command,
# one ncu --set full capture of the hot kernels summarised on the box (the .ncu-rep is too big to bring back).
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -2 gpurun_out/bench_full.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench_1m.csv python bench.py --n-total 1000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none -k regex:"dgemm_kernel|xtfx_kernel|tgemm_kernel|tf32_convert|gemv_rows|splitk_reduce|tgemm_reduce" -c 56 -o /tmp/prof_r01b python tools/profile_kernels.py 500000 1024 1 > gpurun_out/prof.log 2>&1
tail -2 gpurun_out/prof.log
VT_CAPTURE_N=5e5 python tools/summarize_ncu.py /tmp/prof_r01b.ncu-rep gpurun_out/ncu_full_r01b_kernels.csv "ncu --set full --clock-control none; python tools/profile_kernels.py 500000 1024 1; B200, round 1 (after the tcgen05 TF32 engine)" > gpurun_out/summarize.log 2>&1
ncu -i /tmp/prof_r01b.ncu-rep --page raw --csv > gpurun_out/ncu_full_r01b_raw.csv 2>/dev/null
ls -la gpurun_out/ /tmp/prof_r01b.ncu-rep
du -sh gpurun_out
