#!/bin/bash
# round-2 GPU call D: first failing test with traceback, full test run, mgpu N=1, ncu captures of the new kernels
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -60 ) > gpurun_out/r2d_tests_x.log
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2d_tests.log
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 \
    scripts/mgpu_check.py ) > gpurun_out/r2d_mgpu1.log 2>&1
cap() {  # name, kernel regex, config, skip, count, env...
  name=$1; rx=$2; cfg=$3; sk=$4; cnt=$5; shift 5
  env "$@" timeout 420 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $sk -c $cnt \
    -o gpurun_out/r02_$name -f python bench.py --steps 1 --warmup 1 --rows 1e8 --no-cpu --no-e2e --configs $cfg \
    > gpurun_out/r02_${name}_ncu.log 2>&1
}
cap c3_stream b2_join_stream_kernel C3 8 1
cap c3f_joinagg b2_join_agg_kernel C3f 8 1
cap c5_scatter_warp "b2_part_scatter_warp_kernel|b2_part_hist_kernel" C5 16 2
cap c5_scatter_block "b2_part_scatter_kernel" C5 8 1 B200SQL_SCATTER=block
cap c5_ordered "b2_groupby_dense_kernel" C5 1 1
cap c2z_hh "b2_groupby_dense_hh_kernel" C2z 8 1
cap c4s_star "b2_star_agg_kernel" C4s 12 1
ls -la gpurun_out | grep r02_
tail -3 gpurun_out/r2d_tests.log
grep -v "^\[W\|^W0" gpurun_out/r2d_mgpu1.log | tail -3
