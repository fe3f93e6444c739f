set -x
python -m pytest tests/test_gpu_ozaki.py -q -x 2>&1 | tail -3
VT_OGEMM_TIMING=1 python tools/syrk_probe.py 2000000 1024 8000000 320 4000000 512 > gpurun_out/syrk_probe_timing.jsonl
VT_EPI_HELPS=0 VT_OGEMM_TIMING=1 python tools/syrk_probe.py 2000000 1024 8000000 320 > gpurun_out/syrk_probe_nohelp.jsonl
cat gpurun_out/syrk_probe_timing.jsonl gpurun_out/syrk_probe_nohelp.jsonl
VT_OGEMM_TIMING=1 python tools/ogemm_probe.py timing 2>&1 | tail -2
