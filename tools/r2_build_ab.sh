#!/bin/bash
# GPU suite, then A/B of the two build kernels (TCAMCRF_BUILD_DEDUP=0/1/auto) on natural and noise frames.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/ab_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/ab_pytest_gpu.log
tail -4 $O/ab_pytest_gpu.log
rm -f $O/bab_*.json
B="python bench.py --no-cpu-baseline --no-e2e --no-extra --steps 100"
for mode in 0 1 auto; do
  if [ $mode = auto ]; then unset TCAMCRF_BUILD_DEDUP; else export TCAMCRF_BUILD_DEDUP=$mode; fi
  $B --classes 2 --kind natural > $O/bab_${mode}_natural_k2.json 2>> $O/bab.err
  $B --classes 10 --kind natural > $O/bab_${mode}_natural_k10.json 2>> $O/bab.err
  $B --classes 10 > $O/bab_${mode}_noise_k10.json 2>> $O/bab.err
  $B --classes 2 > $O/bab_${mode}_noise_k2.json 2>> $O/bab.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/bab_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = {k: round(v['ms_per_step'], 4) for k, v in d['roofline']['stages'].items()}
        print(f"{f}: fps={d['value']:.0f} ms={d['ms_per_step']:.4f} {st}")
    except Exception as e:
        print(f, 'unreadable', e)
PY
tail -5 $O/bab.err
