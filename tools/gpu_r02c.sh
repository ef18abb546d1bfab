#!/bin/bash
# round 2, call c: row-map force kernel: parity subset, then kernel timing (variant 0 = default, 1 = one-thread-per-point)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "elastic_force_kernel or time_loop_seismograms or mesh_coloring or partitioned or full_size" 2>&1 | grep -v Warning | tail -5
SPECFEM_B200_VERBOSE=1 timeout 900 python tools/bench_force_variants.py --variants 0,1 2>&1 | grep -v Warning | tee gpurun_out/r02c_force_row.jsonl
