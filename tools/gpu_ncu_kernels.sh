#!/bin/bash
# ncu --set full captures of the three dominant force kernels (one launch each) + summaries; see profiles/*_ncu_summary.txt
#   tools/gpu_ncu_kernels.sh <tag>       e.g. r02y -> gpurun_out/r02y_elastic_rownp.ncu-rep, ...
TAG=${1:-ncu}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_elastic_forces_rownp -s 3 -c 1 -o gpurun_out/${TAG}_elastic_rownp -f \
   python tools/bench_force_variants.py --variants 0 --reps 2 > gpurun_out/${TAG}_ncu_el.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_elastic_forces_att_row -s 8 -c 1 -o gpurun_out/${TAG}_att_row -f \
   python bench.py --config cfg3 --nex 96x96x48 --steps 4 --warmup 3 > gpurun_out/${TAG}_ncu_att.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_acoustic_forces_row -s 8 -c 1 -o gpurun_out/${TAG}_acoustic_row -f \
   python tools/bench_acoustic_variants.py --variants 0 > gpurun_out/${TAG}_ncu_ac.log 2>&1; echo "ncu rc=$?"
for k in elastic_rownp att_row acoustic_row; do
  ncu -i gpurun_out/${TAG}_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_${k}_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/${TAG}_${k}_raw.csv > gpurun_out/${TAG}_${k}_ncu_summary.txt
done
