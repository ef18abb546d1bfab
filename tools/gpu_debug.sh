#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tests/ref_cuda_runner.py ocean_source_in_solid /tmp/o.npz > gpurun_out/dbg_ref_ocean.log 2>&1
grep -v "^=========     Host Frame\|^=========         in " gpurun_out/dbg_ref_ocean.log | head -60
python - <<'PY'
import numpy as np
d=np.load('/tmp/o.npz')
for k in d.files: print(k, d[k].shape, float(np.abs(d[k]).max()))
PY
