set -x
for ns in 0 100 500 2000; do
  VT_EPI_SLEEP_NS=$ns VT_OGEMM_TIMING=1 python tools/syrk_probe.py 2000000 1024 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('sleep', $ns, 'syrk_ms', round(d['f64_ozaki_ms'], 2), d['timing'])"
  VT_EPI_SLEEP_NS=$ns VT_OGEMM_TIMING=1 python tools/ogemm_probe.py timing 2>&1 | tail -2
done
ncu --set full --import-source on --clock-control none -k regex:ogemm_kernel --launch-skip 4 --launch-count 1 -o gpurun_out/syrk_fused_r03d python tools/syrk_probe.py 400000 1024 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
