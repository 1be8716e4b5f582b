#!/bin/bash
# GPU suite on the in-tree library, then A/B of prebuilt variants (tools/ab_variants.sh).
#   gpurun --timeout 1200 -- 'bash tools/r2_ab.sh "<kind:K ...>" v1 v2 ...'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/ab_pytest_gpu.log
tail -4 gpurun_out/ab_pytest_gpu.log
bash tools/ab_variants.sh "$@"
