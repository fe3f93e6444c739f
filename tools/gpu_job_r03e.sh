set -x
python -m pytest tests/test_gpu_ozaki.py -q -x 2>&1 | tail -15
VT_OGEMM_TIMING=1 python tools/syrk_probe.py 2000000 1024 8000000 320 > gpurun_out/syrk_probe_timing.jsonl
VT_OZAKI_FUSE=0 python tools/syrk_probe.py 2000000 1024 > gpurun_out/syrk_probe_unfused.jsonl
cat gpurun_out/syrk_probe_timing.jsonl gpurun_out/syrk_probe_unfused.jsonl
VT_OGEMM_TIMING=1 python tools/ogemm_probe.py timing 2>&1 | tail -2
