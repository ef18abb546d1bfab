#!/bin/bash
# round 2, call j: ncu --set full of the viscoelastic row kernel, the acoustic row kernel and the elastic row kernel
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_elastic_forces_att_row -s 8 -c 1 -o gpurun_out/r02j_att_row -f \
   python bench.py --config cfg3 --nex 96x96x48 --steps 4 --warmup 3 > gpurun_out/r02j_ncu_att.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_acoustic_forces_row -s 8 -c 1 -o gpurun_out/r02j_acoustic_row -f \
   python tools/bench_acoustic_variants.py --variants 0 > gpurun_out/r02j_ncu_ac.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_elastic_forces_row -s 3 -c 1 -o gpurun_out/r02j_elastic_row -f \
   python tools/bench_force_variants.py --variants 0 --reps 2 > gpurun_out/r02j_ncu_el.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
