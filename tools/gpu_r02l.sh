#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_elastic_forces_att_row -s 8 -c 1 -o gpurun_out/r02l_att_row -f \
   python bench.py --config cfg3 --nex 96x96x48 --steps 4 --warmup 3 > gpurun_out/r02l_ncu_att.log 2>&1; echo "ncu rc=$?"
