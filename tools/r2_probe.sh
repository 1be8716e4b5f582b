#!/bin/bash
# round-2 first GPU call: GPU suite on the round-1 tree, one bench line, and a probe of compute-sanitizer.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/r2_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_probe_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2_probe_pytest.log
tail -3 $O/r2_probe_pytest.log
timeout 600 python bench.py --no-cpu-baseline > $O/r2_probe_bench.json 2> $O/r2_probe_bench.err
tail -c 400 $O/r2_probe_bench.json
timeout 600 python bench.py --no-cpu-baseline --kind natural --classes 2 --no-e2e > $O/r2_probe_bench_nat2.json 2>> $O/r2_probe_bench.err
# compute-sanitizer probe: one small parity test under memcheck
timeout 600 compute-sanitizer --tool memcheck --log-file $O/r2_sanitizer_memcheck_probe.log \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_lattice_structure_bit_exact and 32" > $O/r2_sanitizer_probe.out 2>&1
echo "sanitizer rc=$?" >> $O/r2_sanitizer_probe.out
tail -5 $O/r2_sanitizer_probe.out
tail -5 $O/r2_sanitizer_memcheck_probe.log
