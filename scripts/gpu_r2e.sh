#!/bin/bash
# multi-GPU call: parity (mgpu_check) + bench at N = number of visible GPUs
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=$(python -c "import torch; print(torch.cuda.device_count())")
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    scripts/mgpu_check.py ) > gpurun_out/r2e_mgpu$N.log 2>&1
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r2e_bench$N.json 2> gpurun_out/r2e_bench$N.err
grep -v "^\[W\|^W0" gpurun_out/r2e_mgpu$N.log | tail -5
cat gpurun_out/r2e_bench$N.json | head -c 3000
