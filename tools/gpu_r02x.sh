#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v Warning | tail -8 | tee gpurun_out/r02x_pytest_gpu_tail.log
for v in 0 1; do
  SPECFEM_B200_KL_KERNEL=$v SPECFEM_B200_VERBOSE=1 timeout 600 python tools/bench_ref_cuda.py --kernel-run --skip-reference --steps 20 --warmup 4 2> gpurun_out/r02x_kl_v$v.err | tail -1 > gpurun_out/r02x_kernel_run_kl_v$v.json
  grep "CTAs per SM" gpurun_out/r02x_kl_v$v.err | sort | uniq -c
  cat gpurun_out/r02x_kernel_run_kl_v$v.json
done
timeout 900 python bench.py 2> gpurun_out/r02x_bench.err | tail -1 > gpurun_out/r02x_bench_n1.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02x_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['newmark'])
for s in d.get('secondary',[]): print(s['row'], s['kernel'], s['kernel_ms'], s['frac'])
print(d.get('config1',{}).get('ms_per_step'), d.get('ref_cuda_backend',{}).get('ms_per_step'))
PY
