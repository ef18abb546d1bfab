#!/bin/bash
# round 2, call i: viscoelastic row kernel: parity, then timing against the one-thread-per-point kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference_cuda.py tests/test_gpu_golden.py -m gpu -x -q -k "attenuation or viscoelastic or cfg3" 2>&1 | grep -v Warning | tail -8
SPECFEM_B200_VERBOSE=1 timeout 900 python bench.py --config cfg3 --nex 96x96x48 --steps 12 --warmup 3 2> gpurun_out/r02i_cfg3_row.err | tee gpurun_out/r02i_cfg3_small_row.json | cut -c1-1800
grep CTAs gpurun_out/r02i_cfg3_row.err
SPECFEM_B200_FORCE_KERNEL=1 timeout 900 python bench.py --config cfg3 --nex 96x96x48 --steps 12 --warmup 3 2>/dev/null | tee gpurun_out/r02i_cfg3_small_pt.json | cut -c1-1800
