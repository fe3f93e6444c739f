# Round-1 final measurements (run under gpurun).  The .ncu-rep stays on the box (too big); its summaries come back.
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ozaki.py -x -q 2>&1 | tail -2
timeout 150 python tools/ogemm_probe.py time_parts > gpurun_out/ogemm_parts_final.jsonl 2>&1
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -3 gpurun_out/bench_full.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench_1m.csv python bench.py --n-total 1000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none -k regex:"dgemm_kernel|xtfx_kernel|tgemm_kernel|tf32_convert|gemv_rows|ogemm_kernel|ozaki_" -c 60 -o /tmp/prof_r01d python tools/profile_kernels.py 200000 1024 1 > gpurun_out/prof.log 2>&1
tail -2 gpurun_out/prof.log
VT_CAPTURE_N=2e5 python tools/summarize_ncu.py /tmp/prof_r01d.ncu-rep gpurun_out/ncu_full_r01d_kernels.csv "ncu --set full --clock-control none; python tools/profile_kernels.py 200000 1024 1; B200, round 1 final (DMMA, TF32 and INT8 engines)" > gpurun_out/summarize.log 2>&1
ncu -i /tmp/prof_r01d.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
keep=[i for i,h in enumerate(hdr) if h in ('Kernel Name','gpu__time_duration.sum') or any(k in h for k in ('pipe_tc','pipe_tensor','dram__bytes','lts__t_bytes.sum','xbar2l1tex','registers_per_thread','tmem','utc'))]
w=csv.writer(sys.stdout)
for r in rows: w.writerow([r[i] for i in keep])
" > gpurun_out/ncu_full_r01d_tensor_metrics.csv
ls -la gpurun_out | head -40; du -sh gpurun_out
