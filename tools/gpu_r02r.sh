#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/bench_ref_cuda.py --kernel-run --steps 20 --warmup 4 --skip-reference 2> gpurun_out/r02r.err | tee gpurun_out/r02r_kernel_run_row.json
SPECFEM_B200_FORCE_KERNEL=1 timeout 1200 python tools/bench_ref_cuda.py --kernel-run --steps 20 --warmup 4 --skip-reference 2>> gpurun_out/r02r.err | tee gpurun_out/r02r_kernel_run_pair_pt.json
timeout 1200 python tools/bench_ref_cuda.py --kernel-run --steps 20 --warmup 4 --only-reference 2> gpurun_out/r02r_ref.err | tee gpurun_out/r02r_kernel_run_reference.json; tail -3 gpurun_out/r02r_ref.err | cut -c1-300
