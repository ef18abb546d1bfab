#!/bin/bash
mkdir -p gpurun_out
SPECFEM_B200_VERBOSE=1 SPECFEM_B200_KL_VARIANT=1 timeout 1200 python tools/bench_ref_cuda.py --kernel-run --steps 20 --warmup 4 --skip-reference 2> gpurun_out/r02s.err | tee gpurun_out/r02s_kernel_run_row_v1.json | cut -c1-400
grep CTAs gpurun_out/r02s.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_elastic_forces_kl_row -s 6 -c 1 -o gpurun_out/r02s_kl_row -f \
   python tools/bench_ref_cuda.py --kernel-run --steps 6 --warmup 2 --skip-reference > gpurun_out/r02s_ncu.log 2>&1; echo "ncu rc=$?"
