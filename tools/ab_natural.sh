#!/bin/bash
# Quick A/B on natural frames: GPU tests, then per-pass blur against the one-launch blur (K = 2 and 10).
#   gpurun --timeout 600 -- 'bash tools/ab_natural.sh'
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/ab_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/ab_pytest_gpu.log
tail -15 $O/ab_pytest_gpu.log
B="python bench.py --no-cpu-baseline --no-e2e --no-extra --steps 200"
rm -f $O/ab_*.json
for k in 2 10; do
  TCAMCRF_BLUR_FRAMES=0 $B --kind natural --classes $k > $O/ab_natural_k${k}_perpass.json 2>> $O/ab.err
  $B --kind natural --classes $k > $O/ab_natural_k${k}_auto.json 2>> $O/ab.err
done
TCAMCRF_BLUR_FRAMES=1 $B --classes 2 > $O/ab_noise_k2_forced.json 2>> $O/ab.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/ab_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = {k: round(v['ms_per_step'], 4) for k, v in d['roofline']['stages'].items()}
        print(f"{f}: fps={d['value']:.0f} ms={d['ms_per_step']:.4f} launches={d.get('gpu_launches')} {st}")
    except Exception as e:
        print(f, 'unreadable', e)
PY
tail -5 $O/ab.err
