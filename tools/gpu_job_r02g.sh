set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py -x -q 2>&1 | tail -3
VT_OGEMM_TIMING=1 timeout 150 python tools/ogemm_probe.py timing > gpurun_out/ogemm_timing_r02g.jsonl 2>&1
cat gpurun_out/ogemm_timing_r02g.jsonl
echo fused; timeout 150 python tools/ogemm_probe.py time_apply
echo serial; VT_OZAKI_FUSE=0 timeout 150 python tools/ogemm_probe.py time_apply
echo overlap; VT_OZAKI_FUSE=0 VT_OZAKI_OVERLAP=1 timeout 150 python tools/ogemm_probe.py time_apply
echo syrk serial; timeout 150 python tools/ogemm_probe.py time_syrk
echo syrk overlap; VT_OZAKI_OVERLAP=1 timeout 150 python tools/ogemm_probe.py time_syrk
