#!/bin/bash
# ncu --set full of seed_fused_kernel (32 samples x 5 frames x 224x224, rng_parity=False)
mkdir -p gpurun_out
O=gpurun_out
ncu --set full --clock-control none --import-source on -k 'regex:seed_fused' --launch-skip 20 --launch-count 1 -f -o $O/r2_seed python tools/seed_timing.py > $O/r2_ncu_seed.log 2>&1
ncu -i $O/r2_seed.ncu-rep --page raw --csv > $O/r2_ncu_seed_raw.csv 2>/dev/null
ncu -i $O/r2_seed.ncu-rep --page source --csv --print-source cuda,sass > $O/r2_ncu_seed_src.csv 2>/dev/null
rm -f $O/r2_seed.ncu-rep
tail -2 $O/r2_ncu_seed.log
