set -x
ncu --set full --import-source on --clock-control none -k regex:ogemm_kernel --launch-skip 4 --launch-count 1 -o gpurun_out/syrk_fused_r03p python tools/syrk_probe.py 400000 1024 > /dev/null 2>&1
ls -la gpurun_out/syrk_fused_r03p.ncu-rep
