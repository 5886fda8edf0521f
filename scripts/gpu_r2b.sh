#!/bin/bash
# round-2 GPU call B: tests (full), mgpu_check N=1 with full log, bench with configs, A/B of join_agg register budgets and scatter kernels
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 ) > gpurun_out/r2b_tests.log
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 \
    scripts/mgpu_check.py ) > gpurun_out/r2b_mgpu1.log 2>&1
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu ) > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
( B200SQL_JA_MINB=2 timeout 300 python bench.py --steps 3 --warmup 3 --rows 1e8 --no-cpu --no-e2e --configs C3f ) \
    > gpurun_out/r2b_c3f_minb2.json 2> gpurun_out/r2b_c3f_minb2.err
( B200SQL_SCATTER=block timeout 300 python bench.py --steps 3 --warmup 3 --rows 1e8 --no-cpu --no-e2e --configs C5 ) \
    > gpurun_out/r2b_c5_block.json 2> gpurun_out/r2b_c5_block.err
( B200SQL_NO_HOT_TABLE=1 timeout 300 python bench.py --steps 3 --warmup 3 --rows 1e8 --no-cpu --no-e2e --configs C2z ) \
    > gpurun_out/r2b_c2z_nohot.json 2> gpurun_out/r2b_c2z_nohot.err
( B200SQL_NO_WARPAGG=1 timeout 300 python bench.py --steps 3 --warmup 3 --rows 1e8 --no-cpu --no-e2e --configs C2z ) \
    > gpurun_out/r2b_c2z_direct.json 2> gpurun_out/r2b_c2z_direct.err
( B200SQL_JOIN_ORDER=counted timeout 300 python bench.py --steps 3 --warmup 3 --rows 1e8 --no-cpu --no-e2e --configs C3 ) \
    > gpurun_out/r2b_c3_counted.json 2> gpurun_out/r2b_c3_counted.err
tail -3 gpurun_out/r2b_tests.log
tail -3 gpurun_out/r2b_mgpu1.log
