#!/bin/bash
# round 2, call p: two-wavefield row kernel for kernel runs: parity (oracle + reference CUDA backend), then step time
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernel_run.py tests/test_gpu_vs_reference_cuda.py -m gpu -x -q -k "kernel" 2>&1 | grep -v Warning | tail -6
SPECFEM_B200_VERBOSE=1 timeout 1200 python tools/bench_ref_cuda.py --kernel-run --steps 20 --warmup 4 2> gpurun_out/r02p.err | tee gpurun_out/r02p_kernel_run.json
grep -E "CTAs" gpurun_out/r02p.err
