#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=$(python -c "import torch; print(torch.cuda.device_count())")
( B200SQL_CALL_TIMES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-e2e ) > gpurun_out/r2g_bench$N.json 2> gpurun_out/r2g_bench$N.err
grep -a 'call times\|select_launch' gpurun_out/r2g_bench$N.err | cut -c1-1500
