#!/bin/bash
# compute-sanitizer memcheck over small invocations of the round's kernels (pytest does not start under the sanitizer here)
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py > gpurun_out/compute_sanitizer_memcheck.log 2>&1
tail -30 gpurun_out/compute_sanitizer_memcheck.log
