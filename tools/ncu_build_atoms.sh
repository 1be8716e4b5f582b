#!/bin/bash
# atomics / loads / time of one build launch for a given set of -D flags
flags="$1"; shift
TCAMCRF_NVCC_EXTRA="$flags" python -c "from tcam_wsol_video_b200 import _lib; _lib.build(force=True)" || exit 1
ncu --metrics l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,lts__t_sectors.sum,gpu__time_duration.sum --clock-control none -k regex:build --launch-skip 4 --launch-count 1 --csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra "$@" 2>/dev/null | grep -E "build_kernel" | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"' | sed "s/^/[$flags] /"
