#!/bin/bash
# kernel-run (SIMULATION_TYPE 3) evidence on one B200: parity tests, same-box bench against the reference CUDA backend
# (fused pair kernel and the stored-strain path), ncu launch list of a few kernel-run steps
mkdir -p gpurun_out
# the reference library malloc()s its Mesh struct and never initialises the flags of features it is not asked to prepare:
# zero-filled heap blocks (glibc: perturb byte 255 -> fill with 0) stand in for the fresh pages it gets under Fortran
export MALLOC_PERTURB_=255
[ -n "$SKIP_TESTS" ] || timeout 600 python -m pytest tests/test_gpu_kernel_run.py "tests/test_gpu_vs_reference_cuda.py::test_against_reference_cuda_backend[kernel_run]" -q 2>&1 \
  | grep -v "Warning\|getlimits\|setattr\|_float_to_str\|^$" | tail -8
[ -n "$SKIP_FUSED" ] || timeout 600 python tools/bench_ref_cuda.py --kernel-run --steps 20 --warmup 3 > gpurun_out/kl_bench_fused.json 2> gpurun_out/kl_bench_fused.err; echo "fused rc=$?"
cat gpurun_out/kl_bench_fused.json
SPECFEM_B200_KL_FUSED=0 timeout 600 python -X faulthandler tools/bench_ref_cuda.py --skip-reference --kernel-run --steps 20 --warmup 3 > gpurun_out/kl_bench_unfused.json 2> gpurun_out/kl_bench_unfused.err; echo "unfused rc=$?"
cat gpurun_out/kl_bench_unfused.json; grep -v Warn gpurun_out/kl_bench_unfused.err | tail -5
[ -n "$SKIP_NCU" ] || timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_kernel_run.csv \
   python tools/bench_ref_cuda.py --kernel-run --steps 3 --warmup 2 > gpurun_out/ncu_kl.log 2>&1; echo "ncu rc=$?"
[ -n "$SKIP_NCU" ] || python tools/summarize_launches.py gpurun_out/launches_kernel_run.csv | head -24
