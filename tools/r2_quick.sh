#!/bin/bash
# GPU suite + short bench lines (stage times): natural K=2 / K=10, noise K=10 / K=2.   $1 = tag
mkdir -p gpurun_out
O=gpurun_out
T=${1:-q}
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest_gpu.log
tail -4 $O/${T}_pytest_gpu.log
rm -f $O/${T}_b_*.json
B="python bench.py --no-cpu-baseline --no-e2e --no-extra --steps 100"
$B --classes 2 --kind natural > $O/${T}_b_natural_k2.json 2>> $O/${T}.err
$B --classes 10 --kind natural > $O/${T}_b_natural_k10.json 2>> $O/${T}.err
$B --classes 10 > $O/${T}_b_noise_k10.json 2>> $O/${T}.err
$B --classes 2 > $O/${T}_b_noise_k2.json 2>> $O/${T}.err
python - $T <<'PY'
import glob, json, sys
for f in sorted(glob.glob(f'gpurun_out/{sys.argv[1]}_b_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = {k: round(v['ms_per_step'], 4) for k, v in d['roofline']['stages'].items()}
        print(f"{f}: fps={d['value']:.0f} ms={d['ms_per_step']:.4f} {st}")
    except Exception as e:
        print(f, 'unreadable', e)
PY
tail -5 $O/${T}.err
