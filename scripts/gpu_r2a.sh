#!/bin/bash
# round-2 GPU call A: correctness of the refactor + micro-benchmarks + all bench configs + scatter A/B
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r2a_tests.log
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 \
    scripts/mgpu_check.py 2>&1 | tail -15 ) > gpurun_out/r2a_mgpu1.log
( timeout 240 ./scripts/microbench/redg ) > gpurun_out/r2a_redg.jsonl 2> gpurun_out/r2a_redg.err
( timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
for v in block warp8 warp16; do
  ( B200SQL_SCATTER=$v timeout 300 python bench.py --steps 3 --warmup 3 --rows 1e8 --no-cpu --no-e2e --configs C5 ) \
    > gpurun_out/r2a_c5_$v.json 2> gpurun_out/r2a_c5_$v.err
done
tail -3 gpurun_out/r2a_tests.log
tail -2 gpurun_out/r2a_mgpu1.log
