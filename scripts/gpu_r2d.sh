#!/bin/bash
# round-2 GPU call D: first failing test with traceback, full test run, mgpu N=1, ncu captures of the new kernels
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
cap() {  # name, kernel regex, config, skip, count, rows per launch, env...
  name=$1; rx=$2; cfg=$3; sk=$4; cnt=$5; rows=$6; shift 6
  env "$@" timeout 420 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $sk -c $cnt \
    -o gpurun_out/r02_$name -f python bench.py --steps 1 --warmup 1 --rows 1e8 --no-cpu --no-e2e --configs $cfg \
    > gpurun_out/r02_${name}_ncu.log 2>&1
  # the report itself is too big to travel back (64 MiB per call): summarise it here, keep the text
  timeout 300 python scripts/ncu_summary.py gpurun_out/r02_$name.ncu-rep $rows > gpurun_out/r02_${name}_summary.txt 2>&1
  rm -f gpurun_out/r02_$name.ncu-rep
}
cap c3_stream b2_join_stream_kernel C3 8 1 125e6
cap c3f_joinagg b2_join_agg_kernel C3f 8 1 125e6
cap c5_scatter_warp "b2_part_scatter_warp_kernel|b2_part_hist_kernel" C5 16 2 62.5e6
cap c5_scatter_block "b2_part_scatter_kernel" C5 8 1 62.5e6 B200SQL_SCATTER=block
cap c5_ordered "b2_groupby_dense_kernel" C5 1 1 500e6
cap c2z_hh "b2_groupby_dense_hh_kernel" C2z 8 1 25e6
cap c4s_star "b2_star_agg_kernel" C4s 12 1 125e6
ls -la gpurun_out | grep r02_
