#!/bin/bash
# round 2, call d: point renumbering + row-map force kernel: whole GPU suite, kernel timing, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_golden.py 2>&1 | grep -v Warning | tail -8 | tee gpurun_out/r02d_pytest_gpu_tail.log
SPECFEM_B200_VERBOSE=1 timeout 900 python tools/bench_force_variants.py --variants 0,1 2>&1 | grep -v Warning | tee gpurun_out/r02d_force_row_renumbered.jsonl
SPECFEM_B200_RENUMBER=0 timeout 900 python tools/bench_force_variants.py --variants 0,1 2>&1 | grep -v Warning | tee gpurun_out/r02d_force_row_hostnumbering.jsonl
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/r02d_bench_n1.json
