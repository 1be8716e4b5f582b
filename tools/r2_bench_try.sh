#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py > $O/try_bench.json 2> $O/try_bench.err
echo "rc=$?"
tail -c 1500 $O/try_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/try_bench.json').read().strip().splitlines()[-1])
r = d['roofline']
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', r['frac'], 'traffic', r['traffic'], r['traffic_what'])
print('parity', d['parity'])
print('cpu', d['cpu_baseline'] and (d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores']))
e = d['e2e']; print('e2e', e['value'], e['ms_per_step'], 'link', e['link_bound_ms'], e['frac_of_link'], e['h2d_gbs'], e['d2h_gbs'], e['h2d_alone_ms'], e['d2h_alone_ms'])
print('e2e_trainer', d['e2e_trainer']['value'], d['e2e_trainer']['float32_frames'])
print('clocks', d['clocks'])
for k, v in d['other_inputs'].items(): print(k, json.dumps(v)[:400])
PY
