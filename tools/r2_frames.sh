#!/bin/bash
# per-frame stage times against the number of frames per step (do the frame tables want to live in L2?)
for n in 8 16 32 64; do
  python bench.py --no-cpu-baseline --no-e2e --no-extra --steps 200 --frames $n --classes 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); n=$n; st={k: round(1e3*v['ms_per_step']/n, 3) for k, v in d['roofline']['stages'].items()}; print('frames=$n us/frame', round(1e3*d['ms_per_step']/n,2), st)"
done
