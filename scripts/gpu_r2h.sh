#!/bin/bash
# one GPU: parity suite, then the full default bench line (all configs, e2e, cpu baseline)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2h_pytest.log 2>&1
tail -3 gpurun_out/r2h_pytest.log
( timeout 900 python bench.py ) > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err
tail -c 600 gpurun_out/r2h_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_bench_n1.json').read().strip().splitlines()[-1])
print('C4 ms/step', d['ms_per_step'], 'value', d['value'], 'ok', d['verified_full_size'])
print('roofline', d['roofline'] and d['roofline']['frac'], 'e2e', d['e2e'] and d['e2e'].get('value'), 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
for k,v in (d['configs'] or {}).items():
    if 'error' in v: print(k, v); continue
    r=v.get('roofline') or {}
    print(k, 'ms %.3f'%v['ms'], 'Grows/s %.1f'%(v['rows_per_s']/1e9), 'whole %.3f'%v['frac_of_peak_whole_query'], 'kernel', r.get('kernel'), 'frac', r.get('frac'), 'ok', v['verified'].get('ok'))
    print('   ', v.get('breakdown_ms_per_step'))
PY
