#!/bin/bash
mkdir -p gpurun_out
for u in 0 1; do
SPECFEM_B200_ATT_UNROLL=$u timeout 900 python bench.py --config cfg3 --nex 96x96x48 --steps 12 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('unroll $u: att kernel ms', r['kernel_ms'], 'frac', r['frac'], 'ms/step', d['ms_per_step'])"
done
