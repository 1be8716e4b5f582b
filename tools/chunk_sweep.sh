#!/bin/bash
# tools/chunk_sweep.sh [bench args...] -- step time against frames per pass (TCAMCRF_CHUNK), library as built
for c in 4 8 16 32; do
TCAMCRF_CHUNK=$c python bench.py --no-cpu-baseline --no-e2e --no-extra --steps 50 "$@" > /tmp/sweep.json 2>/tmp/sweep.err || { tail -5 /tmp/sweep.err; exit 1; }
python - "$c" "$*" <<'PY'
import json, sys
d = json.load(open('/tmp/sweep.json'))
st = {k: round(v['ms_per_step'], 4) for k, v in d['roofline']['stages'].items()}
print(f"[chunk {sys.argv[1]}] [{sys.argv[2]}] fps={d['value']:.0f} ms={d['ms_per_step']:.4f} {st}")
PY
done
