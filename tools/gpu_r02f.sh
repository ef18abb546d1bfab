#!/bin/bash
# round 2, call f: full-length golden runs + the new bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_golden.py -m gpu -q -s > gpurun_out/r02f_gpu_golden_full.log 2>&1
grep -v Warning gpurun_out/r02f_gpu_golden_full.log | grep -E "relative L2|max norms|passed|failed|^E " | cut -c1-1500 | head -40
timeout 1500 python bench.py --steps 50 --warmup 5 > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err
tail -5 gpurun_out/r02f_bench_n1.err
cat gpurun_out/r02f_bench_n1.json
