#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 600 python tools/bench_ref_cuda.py > gpurun_out/bench_ref_cuda.json 2> gpurun_out/bench_ref_cuda.err; echo "refcuda rc=$?"
cat gpurun_out/bench_ref_cuda.json; tail -3 gpurun_out/bench_ref_cuda.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "refarm rc=$?"
cat gpurun_out/bench_reference_arm.json
