#!/bin/bash
# N GPUs of one box: NCCL / NVLink-peer parity check, then the bench with the peer merge and with ncclReduceScatter
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=$(python -c "import torch; print(torch.cuda.device_count())")
( timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    scripts/mgpu_check.py ) > gpurun_out/r2j_check_$N.log 2>&1
echo "check rc=$?"; grep -a "mgpu_check\|Error\|error\|assert\|warn" gpurun_out/r2j_check_$N.log | head -20
run() {
  name=$1; n=$2; shift; shift
  ( env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $n --steps 20 --warmup 5 --no-e2e ) > gpurun_out/r2j_${name}_$n.json 2> gpurun_out/r2j_${name}_$n.err
  echo "== $name n=$n rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2j_${name}_$n.json').read().strip().splitlines()[-1])
    print('ms/step %.3f'%d['ms_per_step'], 'G rows/s %.1f'%(d['value']/1e9), 'ok', d['verified_full_size'].get('ok'), 'groups', d['verified_full_size'].get('groups'), d['verified_full_size'].get('groups_expected'))
    print({k:round(v,3) for k,v in (d['exchange'] or {}).items() if k.endswith('_ms')})
    print('kernel', d['roofline']['avg_launch_ms'], 'share', round(d['roofline']['kernel_share_of_step'],3), (d.get('merge') or '')[:60])
except Exception as e:
    print('no result', e)
PY
  grep -a "warn\|Error" gpurun_out/r2j_${name}_$n.err | head -5
}
for n in $RUN_NS; do
  run peer $n B200SQL_X=1
  run nccl $n B200SQL_PEER_MERGE=0
done
