# Round 2, job d: pair-magic epilogue, converter warps (fused slicing of the next chunk) in the apply.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py -x -q 2>&1 | tail -5
VT_OGEMM_TIMING=1 timeout 150 python tools/ogemm_probe.py timing > gpurun_out/ogemm_timing_r02d.jsonl 2>&1
cat gpurun_out/ogemm_timing_r02d.jsonl
timeout 150 python tools/ogemm_probe.py time_parts > gpurun_out/ogemm_parts_r02d.jsonl 2>&1
timeout 150 python tools/ogemm_probe.py time_apply >> gpurun_out/ogemm_parts_r02d.jsonl 2>&1
VT_OZAKI_FUSE=0 timeout 150 python tools/ogemm_probe.py time_apply >> gpurun_out/ogemm_parts_r02d.jsonl 2>&1
timeout 150 python tools/ogemm_probe.py time_syrk >> gpurun_out/ogemm_parts_r02d.jsonl 2>&1
cat gpurun_out/ogemm_parts_r02d.jsonl
