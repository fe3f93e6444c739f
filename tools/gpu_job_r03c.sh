set -x
python tools/syrk_probe.py 2000000 1024 8000000 320 > gpurun_out/syrk_probe_fused.jsonl
VT_OZAKI_FUSE=0 python tools/syrk_probe.py 2000000 1024 8000000 320 > gpurun_out/syrk_probe_unfused.jsonl
VT_OGEMM_TIMING=1 python tools/syrk_probe.py 2000000 1024 8000000 320 > gpurun_out/syrk_probe_timing.jsonl
cat gpurun_out/syrk_probe_fused.jsonl gpurun_out/syrk_probe_unfused.jsonl gpurun_out/syrk_probe_timing.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/syrk_launches.csv python tools/syrk_probe.py 400000 1024 > /dev/null 2>&1
