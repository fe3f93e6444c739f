# Round 2, final build: launch list of the bench command (share of each kernel in the step) and one ncu --set full
# capture of the hot kernels of the default (INT8 slicing) path.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench_1m_r02z.csv python bench.py --n-total 1000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-tf32 > gpurun_out/bench_under_ncu_r02z.log 2>&1
tail -2 gpurun_out/bench_under_ncu_r02z.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:"ogemm_kernel|ozaki_|xtfx_kernel|chol_diag" -c 24 -o /tmp/prof_r02z python tools/profile_kernels.py 200000 1024 1 > gpurun_out/prof_p.log 2>&1
tail -2 gpurun_out/prof_p.log
VT_CAPTURE_N=2e5 python tools/summarize_ncu.py /tmp/prof_r02z.ncu-rep gpurun_out/ncu_full_r02z_kernels.csv "ncu --set full --clock-control none; python tools/profile_kernels.py 200000 1024 1; B200, round 2 final build (INT8 slicing engine: stacked-B ogemm, integer converter warps for apply and Hessian; stats pass)" > gpurun_out/summarize_p.log 2>&1
cat gpurun_out/ncu_full_r02z_kernels.csv
ncu -i /tmp/prof_r02z.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
keep=[i for i,h in enumerate(hdr) if h in ('Kernel Name','gpu__time_duration.sum') or any(k in h for k in ('pipe_tc','dram__bytes','lts__t_bytes.sum','xbar2l1tex_read_bytes.sum','registers_per_thread','sm__inst_executed_pipe_tc','data_pipe_tc_wavefronts_mem_shared.sum'))]
w=csv.writer(sys.stdout)
for r in rows: w.writerow([r[i] for i in keep])
" > gpurun_out/ncu_full_r02z_tensor_metrics.csv
cp /tmp/prof_r02z.ncu-rep gpurun_out/
