set -x
ncu --set full --import-source on --clock-control none -k regex:ogemm_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/ogemm_apply_r03h python tools/ogemm_probe.py timing > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
