#!/bin/bash
# round-2 baseline at HEAD: GPU suite, headline bench line, natural K=2 line.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_base_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2_base_pytest.log
tail -3 $O/r2_base_pytest.log
timeout 600 python bench.py > $O/r2_base_bench.json 2> $O/r2_base_bench.err
tail -c 600 $O/r2_base_bench.json
timeout 600 python bench.py --no-cpu-baseline --kind natural --classes 2 --no-extra > $O/r2_base_bench_nat2.json 2>> $O/r2_base_bench.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/r2_base_bench*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = {k: round(v['ms_per_step'], 4) for k, v in d['roofline']['stages'].items()}
        print(f"{f}: fps={d['value']:.0f} ms={d['ms_per_step']:.4f} e2e={d['e2e'] and d['e2e']['value']} {st}")
        print({k: (v.get('value') and round(v['value'])) for k, v in d.get('other_inputs', {}).items()})
    except Exception as e:
        print(f, 'unreadable', e)
PY
