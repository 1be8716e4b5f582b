#!/bin/bash
# One GPU-box job: GPU suite, ncu launch lists + `ncu --set full` raw pages of one step for the four workloads,
# profiles/traffic.json regenerated from them for THIS build, then the bench lines (which read it), smoke, reference arm.
#   gpurun --timeout 2400 -- 'bash tools/r2_final.sh <tag>'       (everything lands in gpurun_out/<tag>_*)
T=${1:-r2}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/${T}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest_gpu.log
tail -3 $O/${T}_pytest_gpu.log
P="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra"
K='regex:^(prepare|build|build_dedup|neighbour|vertex_init|splat|splat_rows|blur|slice|loss_backward|loss_backward_logits)_kernel'
SPECS=""
for a in noise:10 natural:2 noise:2 natural:10; do
  IFS=: read kind k <<< "$a"
  tag=${kind}_k${k}
  # launch list of the short command (per-launch times are cold-cache and serialised: shares only)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_${tag}.csv $P --kind $kind --classes $k > $O/${T}_ncu_launches_${tag}.log 2>&1
  # one step under --set full: the LAST step of the run (the second timed step: 11 kernels of the vertex-count probe + 4 steps of 12 are skipped)
  ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 59 --launch-count 12 -f -o $O/${T}_full_${tag} $P --kind $kind --classes $k > $O/${T}_ncu_full_${tag}.log 2>&1
  ncu -i $O/${T}_full_${tag}.ncu-rep --page raw --csv > $O/${T}_ncu_full_raw_${tag}.csv 2>/dev/null
  rm -f $O/${T}_full_${tag}.ncu-rep      # gpurun_out/ comes back only below 64 MiB: keep the exported pages
  SPECS="$SPECS ${kind}:K${k}:N32=$O/${T}_ncu_full_raw_${tag}.csv"
done
SHA=$(python -c "import bench; print(bench.source_sha())")
python tools/make_traffic.py $SHA "ncu --set full --clock-control none, one step of bench.py --steps 3 --warmup 3 (32 frames x 224x224), $T" $SPECS > $O/${T}_traffic.log 2>&1
cp profiles/traffic.json $O/${T}_traffic.json
tail -5 $O/${T}_traffic.log
# the bench lines (they read the traffic.json written above)
t0=$SECONDS
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
echo "bench rc=$? wall $((SECONDS - t0)) s"
python bench.py --kind natural --classes 2 --no-extra > $O/${T}_bench_natural_k2.json 2>> $O/${T}_bench.err
python bench.py --classes 2 --no-extra > $O/${T}_bench_noise_k2.json 2>> $O/${T}_bench.err
python bench.py --kind natural --no-extra > $O/${T}_bench_natural_k10.json 2>> $O/${T}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/${T}_reference_arm.json 2>> $O/${T}_bench.err
python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1
tail -3 $O/${T}_smoke.log
python - $T <<'PY'
import glob, json, sys
t = sys.argv[1]
for f in sorted(glob.glob(f'gpurun_out/{t}_bench*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d['roofline']
        st = {k: (round(v['ms_per_step'], 4), v.get('frac_dram') and round(v['frac_dram'], 2)) for k, v in r['stages'].items()}
        e = d.get('e2e') or {}
        print(f"{f}: fps={d['value']:.0f} ms={d['ms_per_step']:.4f} frac={r['frac']:.3f} dram={r.get('dram', {}).get('frac')} e2e={e.get('value')} link={e.get('frac_of_link')} trainer={(d.get('e2e_trainer') or {}).get('value')} parity={d.get('parity') and (d['parity']['loss_rel'], d['parity']['grad_rel'])}")
        print('   ', st)
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
ls $O | grep "^${T}_" | head -60
