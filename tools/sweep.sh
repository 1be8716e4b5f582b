#!/bin/bash
# tools/sweep.sh "<nvcc -D flags>" [bench args...] -- rebuilds the library with the flags and prints stage times
flags="$1"; shift
TCAMCRF_NVCC_EXTRA="$flags" python -c "from tcam_wsol_video_b200 import _lib; _lib.build(force=True)" || exit 1
python bench.py --no-cpu-baseline --no-e2e --no-extra --steps 50 "$@" > /tmp/sweep.json 2>/tmp/sweep.err || { tail -5 /tmp/sweep.err; exit 1; }
python - "$flags" "$*" <<'PY'
import json, sys
d = json.load(open('/tmp/sweep.json'))
st = {k: round(v['ms_per_step'], 4) for k, v in d['roofline']['stages'].items()}
print(f"[{sys.argv[1]}] [{sys.argv[2]}] fps={d['value']:.0f} ms={d['ms_per_step']:.4f} {st}")
PY
