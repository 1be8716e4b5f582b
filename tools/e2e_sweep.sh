#!/bin/bash
# e2e (host-pointer path) throughput for several pipeline group counts
for g in 1 2 3 4 6 8 16; do
  TCAMCRF_HOST_GROUPS=$g python bench.py --no-cpu-baseline --steps 20 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('groups=$g', 'e2e', round(d['e2e']['value']), 'ms', round(d['e2e']['ms_per_step'],3), 'value', round(d['value']))"
done
