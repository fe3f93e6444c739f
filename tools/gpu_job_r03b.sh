set -x
python -m pytest tests/test_gpu_ozaki.py -q 2>&1 | tail -3
python tools/syrk_probe.py 2000000 1024 1000003 1000 8000000 320 300000 130 > gpurun_out/syrk_probe_fused.jsonl
VT_OZAKI_FUSE=0 python tools/syrk_probe.py 2000000 1024 1000003 1000 8000000 320 300000 130 > gpurun_out/syrk_probe_unfused.jsonl
cat gpurun_out/syrk_probe_fused.jsonl gpurun_out/syrk_probe_unfused.jsonl
