#!/bin/bash
mkdir -p gpurun_out
export SPECFEM_B200_TEST_FORCE_VARIANTS=6,7,9
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "elastic_force_kernel" 2>&1 | grep -v Warning | tail -3
SPECFEM_B200_VERBOSE=1 timeout 900 python tools/bench_force_variants.py --variants 0,5,6,7,9 2>&1 | grep -v Warning | tee gpurun_out/r02w_force_np_ab.jsonl
timeout 900 python tools/bench_att_variants.py --prefetch 3,4,5,6,8 2>&1 | grep -v Warning | tee gpurun_out/r02w_att_prefetch_ab.jsonl
