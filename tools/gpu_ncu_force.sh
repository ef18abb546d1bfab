#!/bin/bash
# one full ncu capture of the force kernel variant $1 (kernel regex $2) on the bench workload
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -o gpurun_out/prof_force_v$1 -f \
   python tools/bench_force_variants.py --variants $1 --reps 2 > gpurun_out/ncu_force_v$1.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_force_v$1.log
