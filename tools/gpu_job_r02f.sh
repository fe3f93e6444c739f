set -x
mkdir -p gpurun_out
export VT_APPLY_N=200000
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s 6 -c 60 --csv --log-file gpurun_out/launches_apply_fused.csv python tools/ogemm_probe.py apply > gpurun_out/ncu_a.log 2>&1
VT_OZAKI_FUSE=0 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s 6 -c 60 --csv --log-file gpurun_out/launches_apply_unfused.csv python tools/ogemm_probe.py apply > gpurun_out/ncu_b.log 2>&1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_syrk.csv python tools/ogemm_probe.py syrk > gpurun_out/ncu_c.log 2>&1
tail -3 gpurun_out/ncu_a.log
ls -la gpurun_out/*.csv
