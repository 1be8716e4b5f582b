#!/bin/bash
# GPU tests, then row-cooperative slice on / off (noise frames, K = 6, 10, 16).
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/sr_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/sr_pytest_gpu.log
tail -15 $O/sr_pytest_gpu.log
rm -f $O/sr_*.json
B="python bench.py --no-cpu-baseline --no-e2e --no-extra --steps 100"
for k in 10 6 16; do
  $B --classes $k > $O/sr_noise_k${k}_rows.json 2>> $O/sr.err
done
$B --classes 10 --kind natural > $O/sr_natural_k10_rows.json 2>> $O/sr.err
TCAMCRF_DENSE=1 $B --classes 10 --kind natural > $O/sr_natural_k10_forced_dense.json 2>> $O/sr.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/sr_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = {k: round(v['ms_per_step'], 4) for k, v in d['roofline']['stages'].items()}
        print(f"{f}: fps={d['value']:.0f} ms={d['ms_per_step']:.4f} {st}")
    except Exception as e:
        print(f, 'unreadable', e)
PY
tail -5 $O/sr.err
