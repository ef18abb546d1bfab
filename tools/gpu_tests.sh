#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q "$@" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -v "Warning\|warnings\|getlimits\|setattr\|_float_to_str" gpurun_out/pytest_gpu.log | tail -40
