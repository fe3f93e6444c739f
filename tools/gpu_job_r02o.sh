set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"block_potrf|block_trsm|block_solve|tall_|gmm_blocks|chol_diag" -c 14 -o /tmp/prof_r02o python tools/profile_blocks.py 200000 > gpurun_out/prof_blocks.log 2>&1
tail -2 gpurun_out/prof_blocks.log
python tools/summarize_ncu.py /tmp/prof_r02o.ncu-rep gpurun_out/ncu_full_r02o_block_kernels.csv "ncu --set full --clock-control none; python tools/profile_blocks.py 200000; B200, r02: block-arrow (N=2e5, M=19, Dg=320) and dense Cholesky (D=4096) kernels" > gpurun_out/summarize_o.log 2>&1
cat gpurun_out/ncu_full_r02o_block_kernels.csv
cp /tmp/prof_r02o.ncu-rep gpurun_out/
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_blocks_r02o.csv python tools/profile_blocks.py 1000000 > gpurun_out/prof_blocks2.log 2>&1
tail -1 gpurun_out/prof_blocks2.log
