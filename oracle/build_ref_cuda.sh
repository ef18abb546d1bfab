#!/bin/bash
# oracle/build_ref_cuda.sh -- TEST INFRASTRUCTURE, not product.
# Compiles the reference's own CUDA backend (SPECFEM3D src/gpu, unmodified, from the sources where they lie under
# /root/reference) for sm_100a into oracle/_ref/libspecfem_ref_cuda.so.  It exports the same Fortran-callable ABI as
# the product library, so the same harness (specfem3d_b200/time_loop.py) can drive it on the B200 box:
#   * GPU-vs-GPU parity against the reference's real kernels (attenuation, acoustic, C-PML, ... where no Fortran
#     CPU build is possible in this image),
#   * the "reference's existing CUDA backend on the same box" number of BASELINE.json north_star.
# Recipe = the reference's own (src/gpu/rules.mk:168-185, Makefile.in:239-244: `nvcc -dc $(GENCODE)` per file, then
# `nvcc -dlink`), with the two configure-generated headers it needs written into the build directory:
#   config.h     : FC_FUNC_ name mangling of gfortran (lower case + underscore)
# WITH_MPI is left undefined (single-process build; the harness does the exchange itself).
# Nothing is copied into the repository; outputs go to oracle/_ref/ only (git-ignored, travels with gpurun).
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
OBJ="$OUT/obj"
LIB="$OUT/libspecfem_ref_cuda.so"
[ -d "$REF/src/gpu" ] || { echo "no reference tree at $REF"; exit 2; }
if [ -f "$LIB" ] && [ -z "${FORCE:-}" ]; then echo "up to date: $LIB"; exit 0; fi
mkdir -p "$OBJ/setup"
cat > "$OBJ/setup/config.h" <<'EOC'
/* written by oracle/build_ref_cuda.sh: what ./configure would generate for gfortran */
#define FC_FUNC(name,NAME) name ## _
#define FC_FUNC_(name,NAME) name ## _
EOC
NVCC=${NVCC:-nvcc}
GEN="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-O3 -lineinfo -dc $GEN -DUSE_CUDA -DGPU_DEVICE_Hopper -Xcompiler -fPIC -I$OBJ/setup -I$REF/setup -I$REF/src/gpu -I$REF/src/gpu/kernels"
objs=()
pids=()
compile() {  # $1 = source, $2 = extra flags
  local o="$OBJ/$(basename "${1%.*}").o"
  objs+=("$o")
  $NVCC -c "$1" -o "$o" $FLAGS $2 &
  pids+=($!)
  if [ ${#pids[@]} -ge ${JOBS:-8} ]; then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi
}
for f in "$REF"/src/gpu/kernels/*.cu; do
  compile "$f" "-include $REF/src/gpu/mesh_constants_gpu.h"
done
for f in "$REF"/src/gpu/*.cu; do
  compile "$f" ""
done
compile "$REF/src/gpu/save_and_compare_cpu_vs_gpu.c" "-x cu"
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -dlink $GEN -Xcompiler -fPIC -o "$OBJ/cuda_device_obj.o" "${objs[@]}"
$NVCC -shared $GEN -o "$LIB" "${objs[@]}" "$OBJ/cuda_device_obj.o" -lcudart
echo "built $LIB"
