#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference_cuda.py tests/test_gpu_golden.py -m gpu -x -q -k "attenuation or viscoelastic or cfg3" 2>&1 | grep -v Warning | tail -4
SPECFEM_B200_VERBOSE=1 timeout 900 python bench.py --config cfg3 --nex 96x96x48 --steps 12 --warmup 3 2> gpurun_out/r02k_cfg3_row.err | tee gpurun_out/r02k_cfg3_small_row.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('att kernel ms', r['kernel_ms'], 'frac', r['frac'], 'ms/step', d['ms_per_step'], 'pml', d['secondary'][0]['kernel_ms'], d['secondary'][0]['frac'])"
grep CTAs gpurun_out/r02k_cfg3_row.err
