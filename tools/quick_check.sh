#!/bin/bash
# GPU suite + a few short bench lines (stage times) after a host-side change.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/qc_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/qc_pytest_gpu.log
tail -6 $O/qc_pytest_gpu.log
rm -f $O/qc_*.json
B="python bench.py --no-cpu-baseline --no-e2e --no-extra --steps 100"
for k in 10 6 16 2; do $B --classes $k > $O/qc_noise_k${k}.json 2>> $O/qc.err; done
$B --classes 2 --kind natural > $O/qc_natural_k2.json 2>> $O/qc.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/qc_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = {k: round(v['ms_per_step'], 4) for k, v in d['roofline']['stages'].items()}
        print(f"{f}: fps={d['value']:.0f} ms={d['ms_per_step']:.4f} {st}")
    except Exception as e:
        print(f, 'unreadable', e)
PY
tail -5 $O/qc.err
