#!/bin/bash
# round 2, call h: acoustic row kernel: parity, then A/B timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernel_run.py -m gpu -x -q -k "acoustic or coupled or ocean or coloring" 2>&1 | grep -v Warning | tail -6
for r in 96 80 72; do SPECFEM_B200_VERBOSE=1 SPECFEM_B200_ACOUSTIC_REGS=$r timeout 600 python tools/bench_acoustic_variants.py --variants 0 2>&1 | grep -E "^\{|CTAs"; done | tee gpurun_out/r02h_acoustic_ab.jsonl
timeout 600 python tools/bench_acoustic_variants.py --variants 1,10 2>&1 | grep -E "^\{" | tee -a gpurun_out/r02h_acoustic_ab.jsonl
