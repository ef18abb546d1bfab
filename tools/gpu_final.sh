#!/bin/bash
# Final evidence of a round on one B200: smoke, N=1 bench (both arms, wall time of each command), ncu launch list of the bench command.
#   tools/gpu_final.sh <tag>
TAG=${1:-final}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warning | tail -2
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$? wall=$(( $(date +%s) - t0 )) s"
cat gpurun_out/${TAG}_bench_n1.json | cut -c1-1500
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "reference arm rc=$? wall=$(( $(date +%s) - t0 )) s"
cut -c1-400 gpurun_out/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --no-ref-cuda > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv | tee gpurun_out/${TAG}_launches.md | head -16
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv
