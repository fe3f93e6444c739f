# Round 2, job b: MMA cost vs N sweep; FP64-Horner 8-warp epilogue; 256-first N partition.
set -x
mkdir -p gpurun_out
VT_PEAK_SECS=0.3 VT_PEAK_N=16,32,48,64,96,128,160,192,208,224,240,256 timeout 200 python tools/ogemm_probe.py i8_peak > gpurun_out/i8_peak_sweep.jsonl 2>&1
cat gpurun_out/i8_peak_sweep.jsonl
timeout 600 python -m pytest tests/test_gpu_ozaki.py -x -q 2>&1 | tail -5
timeout 150 python tools/ogemm_probe.py time_parts > gpurun_out/ogemm_parts_r02b.jsonl 2>&1
timeout 150 python tools/ogemm_probe.py time_apply >> gpurun_out/ogemm_parts_r02b.jsonl 2>&1
timeout 150 python tools/ogemm_probe.py time_syrk >> gpurun_out/ogemm_parts_r02b.jsonl 2>&1
cat gpurun_out/ogemm_parts_r02b.jsonl
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ogemm_kernel" -c 3 -o /tmp/prof_r02b python tools/profile_kernels.py 70000 1024 1 > gpurun_out/prof.log 2>&1
tail -2 gpurun_out/prof.log
python tools/summarize_ncu.py /tmp/prof_r02b.ncu-rep gpurun_out/ncu_full_r02b_kernels.csv "ncu --set full --clock-control none; python tools/profile_kernels.py 70000 1024 1; B200, r02b ogemm (Horner epilogue, 8 warps)" > gpurun_out/summarize.log 2>&1
tail -5 gpurun_out/summarize.log
cp /tmp/prof_r02b.ncu-rep gpurun_out/
