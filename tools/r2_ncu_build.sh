#!/bin/bash
# ncu --set full of the build kernel(s) only: natural K=2 (build_dedup_kernel) and noise K=10 (build_kernel).
mkdir -p gpurun_out
O=gpurun_out
P="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra"
ncu --set full --clock-control none --import-source on -k 'regex:^build' --launch-skip 5 --launch-count 1 -f -o $O/r2_build_nat $P --kind natural --classes 2 > $O/r2_ncu_build_nat.log 2>&1
ncu -i $O/r2_build_nat.ncu-rep --page raw --csv > $O/r2_ncu_build_nat_raw.csv 2>/dev/null
ncu -i $O/r2_build_nat.ncu-rep --page source --csv --print-source cuda,sass > $O/r2_ncu_build_nat_src.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k 'regex:^build' --launch-skip 5 --launch-count 1 -f -o $O/r2_build_noise $P > $O/r2_ncu_build_noise.log 2>&1
ncu -i $O/r2_build_noise.ncu-rep --page raw --csv > $O/r2_ncu_build_noise_raw.csv 2>/dev/null
ncu -i $O/r2_build_noise.ncu-rep --page source --csv --print-source cuda,sass > $O/r2_ncu_build_noise_src.csv 2>/dev/null
ls -la $O | tail
