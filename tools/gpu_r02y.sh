#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "deferred_corrector or newmark or time_loop_seismograms" 2>&1 | grep -v Warning | tail -3
for cm in 1 0; do
  SPECFEM_B200_COMPACT_MASS=$cm timeout 600 python bench.py --no-secondary --no-ref-cuda --steps 40 --warmup 5 2> gpurun_out/r02y_bench_cm$cm.err | tail -1 > gpurun_out/r02y_bench_cm$cm.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r02y_bench_cm$cm.json'))
print('compact_mass=$cm', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['newmark']['kernel_ms'])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_elastic_forces_rownp -s 3 -c 1 -o gpurun_out/r02y_elastic_rownp -f \
   python tools/bench_force_variants.py --variants 0 --reps 2 > gpurun_out/r02y_ncu_el.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_elastic_forces_att_row -s 8 -c 1 -o gpurun_out/r02y_att_row -f \
   python bench.py --config cfg3 --nex 96x96x48 --steps 4 --warmup 3 > gpurun_out/r02y_ncu_att.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02y_launches.csv \
   python bench.py --no-secondary --no-ref-cuda --steps 2 --warmup 1 > gpurun_out/r02y_launches.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r02y*
