#!/bin/bash
# 8-GPU job: host link ceiling with all ranks busy, NCCL tests, then the scaling lines N = 8 (weak, strong 256) and N = 4
mkdir -p gpurun_out
O=gpurun_out
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29544"
timeout 300 $T --nproc-per-node $N tools/pcie_bw.py > $O/m8_pcie_bw.txt 2> $O/m8_pcie_bw.err
tail -3 $O/m8_pcie_bw.txt
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > $O/m8_pytest.log 2>&1
echo "pytest rc=$?" >> $O/m8_pytest.log
tail -3 $O/m8_pytest.log
timeout 600 $T --nproc-per-node $N bench.py --gpus $N --no-extra > $O/m8_weak.json 2> $O/m8_weak.err
timeout 600 $T --nproc-per-node $N bench.py --gpus $N --no-extra --no-e2e --global-frames 256 > $O/m8_strong256.json 2> $O/m8_strong.err
timeout 600 $T --nproc-per-node $N bench.py --gpus $N --no-extra --no-e2e --reduction global > $O/m8_weak_sync.json 2>> $O/m8_weak.err
timeout 600 $T --nproc-per-node 4 bench.py --gpus 4 --no-extra --no-e2e > $O/m4_weak.json 2> $O/m4_weak.err
timeout 600 $T --nproc-per-node 4 bench.py --gpus 4 --no-extra --no-e2e --global-frames 256 > $O/m4_strong256.json 2>> $O/m4_weak.err
timeout 600 python bench.py --no-extra --no-e2e --no-cpu-baseline > $O/m1_ref.json 2> $O/m1.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/m[148]_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get('e2e') or {}
        print(f, 'n', d['n_gpus'], 'fps', round(d['value']), 'ms', round(d['ms_per_step'], 4), d['scaling'], 'coll_us', d['collective_us'] and round(d['collective_us'], 1),
              'e2e', e.get('value') and round(e['value']), 'trainer', d.get('e2e_trainer') and round(d['e2e_trainer']['value']), 'frames/gpu', d['config']['frames_per_gpu'])
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
