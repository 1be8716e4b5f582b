#!/bin/bash
mkdir -p gpurun_out
L=tcam_wsol_video_b200/csrc
for v in seed512 seed1024; do
  cp $L/variants/$v.so $L/libtcamcrf.so
  echo "== $v"
  timeout 600 python -m pytest tests/test_gpu_seeding.py tests/test_gpu_losses.py -m gpu -x -q 2>&1 | tail -1
  python tools/seed_timing.py 2>&1 | tail -3
  python tools/r2_fuzz_seed.py 20 7 2>&1 | tail -1
done
