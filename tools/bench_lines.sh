#!/bin/bash
# smoke(), the three bench lines kept under profiles/ and the reference arm.
#   gpurun --timeout 900 -- 'bash tools/bench_lines.sh'
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/v9_smoke.log 2>&1; echo "smoke rc=$?" >> $O/v9_smoke.log; tail -4 $O/v9_smoke.log
python bench.py > $O/v9_bench.json 2> $O/v9_bench.err
python bench.py --kind natural --no-cpu-baseline > $O/v9_natural_bench.json 2>> $O/v9_bench.err
python bench.py --classes 2 --no-cpu-baseline > $O/v9_k2_bench.json 2>> $O/v9_bench.err
python bench.py --classes 2 --kind natural --no-cpu-baseline > $O/v9_natural_k2_bench.json 2>> $O/v9_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/v9_reference_arm.json 2>> $O/v9_bench.err
tail -c 400 $O/v9_reference_arm.json; tail -3 $O/v9_bench.err
python - <<'PY'
import json
for f in ['v9_bench','v9_natural_bench','v9_k2_bench','v9_natural_k2_bench']:
    d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "trainer", round(d["e2e"]["trainer_call"]["value"]), "trainer_u8", round(d["e2e"]["trainer_call_u8"]["value"]), 'cpu', d['cpu_baseline'] and (round(d['cpu_baseline']['value'],1), d['cpu_baseline'].get('context')))
PY
