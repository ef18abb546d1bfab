#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/bench_ref_cuda.py --kernel-run --steps 20 --warmup 4 --skip-reference 2> gpurun_out/r02q.err | tee gpurun_out/r02q_kernel_run_product.json
SPECFEM_B200_FORCE_KERNEL=1 timeout 1200 python tools/bench_ref_cuda.py --kernel-run --steps 20 --warmup 4 --skip-reference 2>> gpurun_out/r02q.err | tee gpurun_out/r02q_kernel_run_product_pair_pt.json
timeout 1200 python tools/bench_ref_cuda.py --kernel-run --steps 20 --warmup 4 > gpurun_out/r02q_kernel_run_both.json 2> gpurun_out/r02q_both.err; echo "rc=$?"; tail -5 gpurun_out/r02q_both.err | cut -c1-300; cat gpurun_out/r02q_kernel_run_both.json | cut -c1-800
