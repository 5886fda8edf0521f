#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=$(python -c "import torch; print(torch.cuda.device_count())")
run() {
  name=$1; shift
  ( env B200SQL_CALL_TIMES=1 "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $N --steps 20 --warmup 5 --no-e2e ) > gpurun_out/r2g_${name}_$N.json 2> gpurun_out/r2g_${name}_$N.err
  echo "== $name"; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2g_${name}_$N.json').read().strip().splitlines()[-1])
print('ms/step %.3f'%d['ms_per_step'], 'ok', d['verified_full_size']['ok'], {k:round(v,3) for k,v in (d['exchange'] or {}).items() if k.endswith('_ms')})"
  grep -a 'collective host\|call times' gpurun_out/r2g_${name}_$N.err | cut -c1-500
}
run full
run noprep B200SQL_NO_PREPARED=1
