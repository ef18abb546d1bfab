"""TEST INFRASTRUCTURE, not product: the reference's METIS partitioner (external_libs/metis-4.0.3.tar.gz, compiled unmodified
into oracle/_ref/libmetis_ref.so by oracle/build_ref_metis.sh) called exactly like xdecompose_mesh calls it for
PARTITIONING_TYPE = 2:

    src/decompose_mesh/partition_metis.F90:52-79   vwgt(1,:) = elmnts_load(:), one constraint, no edge weights
    src/decompose_mesh/wrap_metis.c:139-153        METIS_PartGraphKway(nspec, xadj, adjncy, vwgt, NULL, wgtflag = 2,
                                                   numflag = 0, nparts, options = {0,0,0,0,0}, edgecut, part)

Used by tests/test_decompose.py (bit-exact pin of part[]) and tools/make_metis_partition.py (partition files for bench.py)."""
from __future__ import annotations

import ctypes
import os

import numpy as np

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libmetis_ref.so")


def available() -> bool:
    return os.path.exists(LIB)


def partition_metis(xadj: np.ndarray, adjncy: np.ndarray, loads: np.ndarray, nparts: int):
    """(part[nspec] 0-based, edgecut).  nparts == 1 returns zeros like partition_metis.F90:63."""
    nspec = len(xadj) - 1
    if nparts == 1:
        return np.zeros(nspec, np.int32), 0
    lib = ctypes.CDLL(LIB)
    xadj = np.ascontiguousarray(xadj, np.int32)
    adjncy = np.ascontiguousarray(adjncy, np.int32)
    vwgt = np.ascontiguousarray(loads, np.int32)
    part = np.zeros(nspec, np.int32)
    n, wgtflag, numflag, np_, edgecut = (ctypes.c_int(v) for v in (nspec, 2, 0, nparts, 0))
    options = (ctypes.c_int * 5)(0, 0, 0, 0, 0)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.METIS_PartGraphKway(ctypes.byref(n), P(xadj), P(adjncy), P(vwgt), None, ctypes.byref(wgtflag), ctypes.byref(numflag),
                            ctypes.byref(np_), options, ctypes.byref(edgecut), P(part))
    if edgecut.value < 0:
        raise RuntimeError("METIS Error: negative edgecut")
    return part, int(edgecut.value)
