#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_seeding.py tests/test_gpu_losses.py -m gpu -x -q > $O/s_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/s_pytest_gpu.log
tail -30 $O/s_pytest_gpu.log
python tools/seed_timing.py 2>&1 | tail -8
