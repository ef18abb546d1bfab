#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "elastic_force_kernel" 2>&1 | grep -v Warning | tail -3
SPECFEM_B200_VERBOSE=1 timeout 900 python tools/bench_force_variants.py --variants 0,2,3 2>&1 | grep -v Warning | tee gpurun_out/r02e_force_ab_renumbered.jsonl
SPECFEM_B200_RENUMBER=0 timeout 900 python tools/bench_force_variants.py --variants 0,2,3 2>&1 | grep -v Warning | tee gpurun_out/r02e_force_ab_hostnumbering.jsonl
