#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "elastic_force_kernel or time_loop_seismograms" 2>&1 | tail -3
timeout 900 python tools/bench_force_variants.py --variants "$1" 2>&1 | grep -v Warning | tee gpurun_out/ab_variants.jsonl
