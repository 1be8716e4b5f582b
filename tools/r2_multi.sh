#!/bin/bash
# N-GPU job: NCCL tests of the sharding module, then the bench lines (weak scaling and the 256-frame strong scaling)
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > $O/m${N}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/m${N}_pytest.log
tail -3 $O/m${N}_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $T bench.py --gpus $N --no-extra > $O/m${N}_weak.json 2> $O/m${N}_weak.err
timeout 600 $T bench.py --gpus $N --no-extra --global-frames 256 --no-e2e > $O/m${N}_strong256.json 2> $O/m${N}_strong.err
timeout 600 $T bench.py --gpus $N --no-extra --reduction global --no-e2e > $O/m${N}_weak_sync.json 2>> $O/m${N}_weak.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for f in (f'gpurun_out/m{n}_weak.json', f'gpurun_out/m{n}_strong256.json', f'gpurun_out/m{n}_weak_sync.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get('e2e') or {}
        print(f, 'fps', round(d['value']), 'ms', round(d['ms_per_step'], 4), d['scaling'], 'coll_us', d['collective_us'],
              'e2e', e.get('value') and round(e['value']), 'trainer', d.get('e2e_trainer') and round(d['e2e_trainer']['value']),
              d['config']['parallelism'], d['config']['frames_per_gpu'])
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
tail -3 $O/m${N}_weak.err
