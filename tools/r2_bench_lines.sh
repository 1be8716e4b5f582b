#!/bin/bash
# the four bench lines + reference arm only (after a change that does not touch the kernels): gpurun_out/<tag>_bench*.json
T=${1:-r2}
mkdir -p gpurun_out
O=gpurun_out
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
echo "bench rc=$?"
python bench.py --kind natural --classes 2 --no-extra > $O/${T}_bench_natural_k2.json 2>> $O/${T}_bench.err
python bench.py --classes 2 --no-extra > $O/${T}_bench_noise_k2.json 2>> $O/${T}_bench.err
python bench.py --kind natural --no-extra > $O/${T}_bench_natural_k10.json 2>> $O/${T}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/${T}_reference_arm.json 2>> $O/${T}_bench.err
python - $T <<'PY'
import json, sys
t = sys.argv[1]
b = json.loads(open(f'gpurun_out/{t}_bench.json').read().strip().splitlines()[-1])
print(round(b['value']), b['ms_per_step'], b['roofline']['frac'], b['roofline']['traffic'], b['e2e']['value'], b['e2e_trainer']['value'], b['parity'])
for k, v in b['other_inputs'].items():
    print(k, json.dumps(v)[:260])
PY
