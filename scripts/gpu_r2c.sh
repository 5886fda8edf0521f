#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -80 ) > gpurun_out/r2c_tests.log
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 \
    scripts/mgpu_check.py ) > gpurun_out/r2c_mgpu1.log 2>&1
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu ) > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
ab() {  # name, configs, env...
  name=$1; cfg=$2; shift 2
  ( env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --rows 1e8 --no-cpu --no-e2e --configs $cfg ) \
    > gpurun_out/r2c_$name.json 2> gpurun_out/r2c_$name.err
}
ab c2_pipe C2 B200SQL_PIPELINE=1
ab c2z_warp C2z B200SQL_SKEW=warp
ab c3_warpres C3 B200SQL_JOIN_RESERVE=warp
ab c5_block C5 B200SQL_SCATTER=block
tail -3 gpurun_out/r2c_tests.log
grep -v "^\[W\|^W0" gpurun_out/r2c_mgpu1.log | tail -4
