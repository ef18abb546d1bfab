#!/bin/bash
# end-of-round check on one B200: smoke(), all GPU tests, N=1 bench (both arms), ncu launch list + full capture of the force kernel
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "Warning\|getlimits\|setattr\|_float_to_str" | tail -2
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -v "Warning\|warnings\|getlimits\|setattr\|_float_to_str\|^tests/\|^$" gpurun_out/pytest_gpu.log | tail -6
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cut -c1-250 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
cut -c1-200 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
python tools/summarize_launches.py gpurun_out/launches.csv | head -12
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_elastic_forces -s 6 -c 1 -f -o gpurun_out/prof_force \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_force.ncu-rep --page raw --csv > gpurun_out/prof_force_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_force_raw.csv > gpurun_out/prof_force_summary.txt 2>&1; head -8 gpurun_out/prof_force_summary.txt
