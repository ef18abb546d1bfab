#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernel_run.py tests/test_gpu_vs_reference_cuda.py -m gpu -q -s -k "undoatt_visco or run_viscoelastic" 2>&1 | grep -v Warning | grep -v "^$" | grep "undo\|passed\|failed\|Error\|^E " | head -40 | tee gpurun_out/r02z_undoatt_tests.log
