#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/r2f_tests.log
( timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu ) > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
( B200SQL_NO_PREPARED=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-configs ) \
    > gpurun_out/r2f_bench_noprep.json 2> gpurun_out/r2f_bench_noprep.err
tail -3 gpurun_out/r2f_tests.log
