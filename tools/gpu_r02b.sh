#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_golden.py -m gpu -q -s -k "cfg1 or waterlayered" > gpurun_out/r02b_gpu_golden_full.log 2>&1
grep -v Warning gpurun_out/r02b_gpu_golden_full.log | grep -E "relative L2|passed|failed|Error|assert|^E " | head -40
