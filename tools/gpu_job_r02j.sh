set -x
mkdir -p gpurun_out
python bench.py --n-total 1000000 --steps 2 --warmup 1 --e2e-steps 1 > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err
tail -3 gpurun_out/bench_1m.err; cat gpurun_out/bench_1m.json
python bench.py > gpurun_out/bench_full_r02j.json 2> gpurun_out/bench_full_r02j.err
tail -3 gpurun_out/bench_full_r02j.err; cat gpurun_out/bench_full_r02j.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r02j.json 2> gpurun_out/bench_ref_r02j.err
tail -3 gpurun_out/bench_ref_r02j.err; cat gpurun_out/bench_ref_r02j.json
