#!/bin/bash
# ncu --set full of the fused-softmax kernels (natural K=2, from logits): splat<L>, slice<L>, loss_backward_logits
mkdir -p gpurun_out
O=gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:(splat_kernel<5, 2, 1>|slice_kernel<5, 2, 1>|loss_backward_logits_kernel)' --launch-skip 90 --launch-count 3 -f -o $O/r2_logits python tools/r2_logits_stages.py > $O/r2_ncu_logits.log 2>&1
ncu -i $O/r2_logits.ncu-rep --page raw --csv > $O/r2_ncu_logits_raw.csv 2>/dev/null
ncu -i $O/r2_logits.ncu-rep --page source --csv --print-source cuda,sass > $O/r2_ncu_logits_src.csv 2>/dev/null
rm -f $O/r2_logits.ncu-rep
tail -3 $O/r2_ncu_logits.log
