#!/bin/bash
# 8 GPUs of one box: parity check at 8 ranks, bench at N=8 (NVLink peer merge vs ncclReduceScatter), then two N=4 variants side by side
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=$(python -c "import torch; print(torch.cuda.device_count())")
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    scripts/mgpu_check.py ) > gpurun_out/r2k_check_$N.log 2>&1
echo "check rc=$?"; grep -a "mgpu_check OK\|Error\|assert\|UserWarning" gpurun_out/r2k_check_$N.log | head -10
run() {
  name=$1; n=$2; port=$3; devs=$4; shift; shift; shift; shift
  ( env CUDA_VISIBLE_DEVICES=$devs "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 20 --warmup 5 --no-e2e ) > gpurun_out/r2k_${name}_$n.json 2> gpurun_out/r2k_${name}_$n.err
  echo "== $name n=$n rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2k_${name}_$n.json').read().strip().splitlines()[-1])
    print('ms/step %.3f'%d['ms_per_step'], 'G rows/s %.1f'%(d['value']/1e9), 'ok', d['verified_full_size'].get('ok'), 'groups', d['verified_full_size'].get('groups'), d['verified_full_size'].get('groups_expected'))
    print({k:round(v,3) for k,v in (d['exchange'] or {}).items() if k.endswith('_ms')})
    print('kernel', round(d['roofline']['avg_launch_ms'],4), 'share', round(d['roofline']['kernel_share_of_step'],3), (d.get('merge') or '')[:13], 'clocks', d['clocks'])
except Exception as e:
    print('no result', e)
PY
  grep -a "UserWarning\|Error" gpurun_out/r2k_${name}_$n.err | head -3
}
if [ "$N" -ge 8 ]; then
  run peer 8 29512 0,1,2,3,4,5,6,7 B200SQL_X=1
  run nccl 8 29512 0,1,2,3,4,5,6,7 B200SQL_PEER_MERGE=0
  run peer 4 29513 0,1,2,3 B200SQL_X=1 &
  run peerroot 4 29514 4,5,6,7 B200SQL_BENCH_DIM=root &
  wait
else
  run peer $N 29512 $(seq -s, 0 $((N-1))) B200SQL_X=1
  run nccl $N 29512 $(seq -s, 0 $((N-1))) B200SQL_PEER_MERGE=0
  run peerroot $N 29512 $(seq -s, 0 $((N-1))) B200SQL_BENCH_DIM=root
fi
