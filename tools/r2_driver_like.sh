#!/bin/bash
# what the driver runs at round end, with its flags: GPU suite, smoke, reference arm, our arm (N=1, --steps 20 --warmup 3)
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > $O/drv_reference.json 2> $O/drv.err; echo "reference rc=$?"
t0=$SECONDS
python bench.py --gpus 1 --steps 20 --warmup 3 > $O/drv_ours.json 2>> $O/drv.err; echo "ours rc=$? wall $((SECONDS - t0)) s"
python - <<'PY'
import json
r = json.loads(open('gpurun_out/drv_reference.json').read().strip().splitlines()[-1])
o = json.loads(open('gpurun_out/drv_ours.json').read().strip().splitlines()[-1])
print('reference', round(r['value'], 1), r['unit'], r['cpu_baseline']['kind'], r['cpu_baseline']['cores'], 'steps', r['steps'])
print('ours value', round(o['value']), 'e2e', round(o['e2e']['value']), 'ratio', round(o['value'] / r['value'], 1), 'e2e ratio', round(o['e2e']['value'] / r['value'], 1),
      'frac', round(o['roofline']['frac'], 3), 'traffic', o['roofline']['traffic'], 'launches', o['gpu_launches'], 'parity', o['parity']['ok'], 'clocks', o['clocks'])
PY
