set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/chol_launches_r03i.csv python tools/chol_once.py 4096 2048 > /dev/null 2>&1
wc -l gpurun_out/chol_launches_r03i.csv
