#!/bin/bash
# One GPU-box job: GPU test suite, the bench lines, the ncu launch lists and the ncu --set full captures of one
# step (noise K=10, natural K=2).  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/r1_v9_run.sh'
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/v9_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/v9_pytest_gpu.log 2>&1
rc=$?
echo "pytest rc=$rc" >> $O/v9_pytest_gpu.log
tail -5 $O/v9_pytest_gpu.log
# the bench lines (headline + the two other regimes)
python bench.py > $O/v9_bench.json 2> $O/v9_bench.err
python bench.py --kind natural --no-cpu-baseline > $O/v9_natural_bench.json 2>> $O/v9_bench.err
python bench.py --classes 2 --no-cpu-baseline > $O/v9_k2_bench.json 2>> $O/v9_bench.err
tail -c 600 $O/v9_bench.json
# launch list of the same command (short)
P="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/v9_launches.csv $P > $O/v9_ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/v9_launches_natural_k2.csv $P --kind natural --classes 2 > $O/v9_ncu_launches_nat.log 2>&1
# one step under --set full (our kernels only), noise K=10 and natural K=2
ncu --set full --clock-control none --import-source on -k 'regex:^(prepare|build|neighbour|vertex_init|splat|splat_rows|blur|blur_frames|slice|loss_backward)_kernel' --launch-skip 59 --launch-count 12 -f -o $O/v9_full_noise $P > $O/v9_ncu_full_noise.log 2>&1
ncu -i $O/v9_full_noise.ncu-rep --page raw --csv > $O/v9_ncu_full_raw_noise.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k 'regex:^(prepare|build|neighbour|vertex_init|splat|splat_rows|blur|blur_frames|slice|loss_backward)_kernel' --launch-skip 59 --launch-count 12 -f -o $O/v9_full_natural_k2 $P --kind natural --classes 2 > $O/v9_ncu_full_nat.log 2>&1
ncu -i $O/v9_full_natural_k2.ncu-rep --page raw --csv > $O/v9_ncu_full_raw_natural_k2.csv 2>/dev/null
ls -la $O | tail -30
