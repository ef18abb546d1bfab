#!/bin/bash
# compute-sanitizer over the hot-path kernels on one B200: memcheck (out-of-bounds / misaligned global and shared accesses),
# synccheck (barrier / mbarrier misuse) and racecheck (shared-memory hazards) on small parity tests that launch every force
# kernel (row kernels with their TMA rings, point kernels, C-PML), the Newmark passes, Stacey, coupling, sources, seismograms
# and both halo paths.      tools/gpu_sanitizer.sh <tag>
TAG=${1:-sanitizer}
mkdir -p gpurun_out
SEL=${SEL:-"test_elastic_force_kernel_matches_oracle or test_attenuation_force_kernel_matches_oracle or (test_acoustic_force_kernel_matches_oracle and nex0) or test_newmark_stacey_sources_norm_kernels or test_p2p_halo_matches_host_staged_halo or (test_cpml_time_loop_matches_oracle and irregular)"}
for tool in ${TOOLS:-memcheck synccheck racecheck}; do
  t0=$(date +%s)
  timeout ${LIMIT:-120} compute-sanitizer --tool $tool --error-exitcode 7 --log-file gpurun_out/${TAG}_${tool}.log \
      python -m pytest ${FILE:-tests/test_gpu_parity.py} -m gpu -q -x -k "$SEL" > gpurun_out/${TAG}_${tool}_pytest.log 2>&1
  echo "$tool rc=$? wall=$(( $(date +%s) - t0 )) s: $(grep -E 'passed|failed|error' gpurun_out/${TAG}_${tool}_pytest.log | tail -1)"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/${TAG}_${tool}.log | tail -2
  grep -E "Invalid|hazard|Barrier error|divergent" gpurun_out/${TAG}_${tool}.log | sort | uniq -c | sort -rn | head -5
done
