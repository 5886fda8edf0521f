#!/bin/bash
# one GPU: parity suite (new: peer merge on one GPU, AoS star hash), C2 step diagnostics, C5 scatter variants, C4s
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2i_pytest.log 2>&1
tail -4 gpurun_out/r2i_pytest.log
( timeout 300 python scripts/diag_c2_steps.py float ) > gpurun_out/r2i_c2_steps.log 2>&1
grep -a "step" gpurun_out/r2i_c2_steps.log | head -20
summ() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('C4 ms/step', round(d['ms_per_step'],3), 'roofline', d['roofline'] and round(d['roofline']['frac'],3), 'ok', d['verified_full_size']['ok'])
for k,v in (d['configs'] or {}).items():
    if 'error' in v: print(k, v); continue
    r=v.get('roofline') or {}
    print(k, 'ms %.3f'%v['ms'], 'whole %.3f'%v['frac_of_peak_whole_query'], 'kernel', r.get('kernel'), 'frac', r.get('frac') and round(r['frac'],3), 'ok', v['verified'].get('ok'))
    print('   ', v.get('breakdown_ms_per_step'))
PY
}
( timeout 600 python bench.py --rows 2e8 --no-e2e --no-cpu --configs C4s,C5,C2 ) > gpurun_out/r2i_a.json 2> gpurun_out/r2i_a.err
summ gpurun_out/r2i_a.json
( B200SQL_SCATTER=block timeout 600 python bench.py --rows 2e8 --no-e2e --no-cpu --configs C5 ) > gpurun_out/r2i_b.json 2> gpurun_out/r2i_b.err
summ gpurun_out/r2i_b.json
tail -c 300 gpurun_out/r2i_a.err
