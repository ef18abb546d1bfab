#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernel_run.py tests/test_gpu_vs_reference_cuda.py -m gpu -x -q -k "kernel" 2>&1 | grep -v Warning | tail -3
for v in 0 1; do SPECFEM_B200_KL_VARIANT=$v timeout 1200 python tools/bench_ref_cuda.py --kernel-run --steps 20 --warmup 4 --skip-reference 2>/dev/null | tee gpurun_out/r02t_kernel_run_row_v$v.json | cut -c1-330; done
