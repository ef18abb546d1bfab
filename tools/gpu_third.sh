#!/bin/bash
mkdir -p gpurun_out
export SPECFEM_B200_CHECK=1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "elastic_force_kernel" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -15 gpurun_out/sanitizer_memcheck.log
unset SPECFEM_B200_CHECK
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_elastic_forces -s 6 -c 2 -o gpurun_out/prof_force_grp \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
