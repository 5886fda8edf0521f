#!/bin/bash
# Round evidence for profiles/: (1) launch list of the bench command, (2) full ncu capture of the
# dominant kernel inside the same command.  Run under gpurun (1 GPU); numbers printed by bench.py
# under ncu are NOT bench values.
set -x
mkdir -p gpurun_out
R=${1:-r01}
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${R}_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${R}_bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:b2_star_agg_kernel -s 8 -c 2 \
    -o gpurun_out/${R}_star python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${R}_star_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:"b2_(scan_agg|groupby_dense|join_onepass|join_count8|select_write)_kernel" -c 8 \
    -o gpurun_out/${R}_others python scripts/ncu_target2.py > gpurun_out/${R}_others_ncu.log 2>&1
ls -la gpurun_out
