#!/bin/bash
# oracle/build_ref_metis.sh -- TEST INFRASTRUCTURE, not product.
# Builds the METIS 4.0.3 the reference vendors (external_libs/metis-4.0.3.tar.gz, plain C) into
# oracle/_ref/libmetis_ref.so, so that the partition `xdecompose_mesh` computes with PARTITIONING_TYPE = 2
# (src/decompose_mesh/partition_metis.F90:52-79 -> wrap_metis.c:95-160: METIS_PartGraphKway, vertex weights = element loads,
# no edge weights, C numbering, default options) can be reproduced bit for bit: tests/test_decompose.py pins `part[]`.
# The tarball is unpacked into a scratch directory under oracle/_ref/ (git-ignored), compiled with gcc, and the scratch
# directory is removed again; nothing of METIS is copied into the repository.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
LIB="$OUT/libmetis_ref.so"
TAR="$REF/external_libs/metis-4.0.3.tar.gz"
[ -f "$TAR" ] || { echo "no METIS tarball at $TAR"; exit 2; }
if [ -f "$LIB" ] && [ -z "${FORCE:-}" ]; then echo "up to date: $LIB"; exit 0; fi
mkdir -p "$OUT/metis_build"
tar xzf "$TAR" -C "$OUT/metis_build"
SRC="$OUT/metis_build/metis-4.0.3/Lib"
# 1998 C: implicit ints / K&R declarations; its own `int log2(int)` clashes with <math.h>'s, so the system headers are
# pulled in first and METIS's function is renamed for the rest of every translation unit (a build flag, no source edit)
printf '#include <math.h>\n#include <stdlib.h>\n#include <string.h>\n#define log2 metis403_log2\n' > "$OUT/metis_build/prefix.h"
gcc -O2 -fPIC -shared -std=gnu89 -w -fcommon -include "$OUT/metis_build/prefix.h" -I"$SRC" -o "$LIB" "$SRC"/*.c -lm
rm -rf "$OUT/metis_build"
echo "built $LIB"
