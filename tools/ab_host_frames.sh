#!/bin/bash
# GPU tests, then the trainer-side call (pinned host frames, device segmentations) for several section counts.
#   gpurun --timeout 600 -- 'bash tools/ab_host_frames.sh'
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/hf_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/hf_pytest_gpu.log
tail -15 $O/hf_pytest_gpu.log
rm -f $O/hf_*.json
for k in 10 2; do for kind in noise natural; do for s in 1 2 4 8; do
  TCAMCRF_HIMG_SECTIONS=$s python bench.py --no-cpu-baseline --no-extra --steps 30 --e2e-steps 40 --classes $k --kind $kind > $O/hf_${kind}_k${k}_s${s}.json 2>> $O/hf.err
done; done; done
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/hf_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        t = d['e2e']['trainer_call']
        print(f"{f}: device-resident ms={d['ms_per_step']:.4f}  trainer_call ms={t['ms_per_step']:.4f} fps={t['value']:.0f}  e2e fps={d['e2e']['value']:.0f}")
    except Exception as e:
        print(f, 'unreadable', e)
PY
tail -5 $O/hf.err
