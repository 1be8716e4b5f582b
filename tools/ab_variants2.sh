#!/bin/bash
# A/B of prebuilt library variants with the GPU suite run on the LAST one:  bash tools/ab_variants2.sh "<kind:K ...>" v1 v2 ...
mkdir -p gpurun_out
O=gpurun_out
L=tcam_wsol_video_b200/csrc
args="$1"; shift
rm -f $O/var_*.json
for rep in 1 2; do
for v in "$@"; do
  cp $L/variants/$v.so $L/libtcamcrf.so
  for a in $args; do
    IFS=: read kind k <<< "$a"
    python bench.py --no-cpu-baseline --no-e2e --no-extra --steps 300 --kind $kind --classes $k > $O/var_${v}_${kind}_k${k}_r${rep}.json 2>> $O/var.err
  done
done
done
timeout 900 python -m pytest tests -m gpu -x -q > $O/var_pytest_gpu.log 2>&1
echo "pytest rc=$? (on the last variant)"; tail -2 $O/var_pytest_gpu.log
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/var_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = {k: round(v['ms_per_step'], 4) for k, v in d['roofline']['stages'].items()}
        print(f"{f}: fps={d['value']:.0f} ms={d['ms_per_step']:.4f} {st}")
    except Exception as e:
        print(f, 'unreadable', e)
PY
tail -3 $O/var.err
