#!/bin/bash
# round 2, call a: layout micro-benchmark + full-length golden runs on the GPU
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_gpu.txt
timeout 300 tools/micro/bin/gather_red_layouts > gpurun_out/r02a_gather_red_layouts.jsonl 2> gpurun_out/r02a_micro.err
cat gpurun_out/r02a_gather_red_layouts.jsonl
timeout 2400 python -m pytest tests/test_gpu_golden.py -m gpu -q -s 2>&1 | grep -v Warning | tail -40 | tee gpurun_out/r02a_gpu_golden.log
