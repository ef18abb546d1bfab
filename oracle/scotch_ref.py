"""TEST INFRASTRUCTURE, not product: the reference's SCOTCH partitioner (external_libs/scotch_5.1.12b, compiled unmodified into
oracle/_ref/libscotch_ref.so by oracle/build_ref_scotch.sh) called exactly like xdecompose_mesh calls it for
PARTITIONING_TYPE = 1 (its default) -- through the library's own Fortran entry points, with the argument aliasing of the call site:

    src/decompose_mesh/partition_scotch.F90:124   call scotchfstratinit(scotchstrat, ier)          empty = default strategy
    src/decompose_mesh/partition_scotch.F90:130   call scotchfrandomReset()                        deterministic partitions
    src/decompose_mesh/partition_scotch.F90:137   call scotchfgraphinit(scotchgraph, ier)
    src/decompose_mesh/partition_scotch.F90:163   call scotchfgraphbuild(scotchgraph, 0, nspec, xadj, xadj, elmnts_load, xadj,
                                                                         nb_edges, adjncy, adjncy, ier)
                                                  (vendtab = verttab: compact arrays; vlbltab = verttab: no labels;
                                                   edlotab = edgetab: no edge loads -- library_graph.c:233-239)
    src/decompose_mesh/partition_scotch.F90:172   call scotchfgraphcheck(scotchgraph, ier)
    src/decompose_mesh/partition_scotch.F90:177   call scotchfgraphpart(scotchgraph, nparts, scotchstrat, part, ier)
    src/decompose_mesh/partition_scotch.F90:184   call scotchfgraphexit / scotchfstratexit

Used by tests/test_decompose.py (bit-exact pin of part[]) and tools/make_metis_partition.py --partitioner scotch."""
from __future__ import annotations

import ctypes
import os

import numpy as np

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libscotch_ref.so")


def available() -> bool:
    return os.path.exists(LIB)


def partition_scotch(xadj: np.ndarray, adjncy: np.ndarray, loads: np.ndarray, nparts: int) -> np.ndarray:
    """part[nspec], 0-based.  nparts == 1 returns zeros (decompose_mesh.F90:127-128 never calls the partitioner then)."""
    nspec = len(xadj) - 1
    if nparts <= 1:
        return np.zeros(nspec, np.int32)
    lib = ctypes.CDLL(LIB)
    xadj = np.ascontiguousarray(xadj, np.int32)
    adjncy = np.ascontiguousarray(adjncy, np.int32)
    load = np.ascontiguousarray(loads, np.int32)
    part = np.full(nspec, -1, np.int32)
    # scotchf.h: SCOTCH_GRAPHDIM = 12, SCOTCH_STRATDIM = 1 doubles on this build; opaque, so simply generous
    graph = np.zeros(64, np.float64)
    strat = np.zeros(64, np.float64)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    I = lambda v: ctypes.byref(ctypes.c_int(v))
    ier = ctypes.c_int(0)

    def check(what):
        if ier.value != 0:
            raise RuntimeError(f"Scotch ERROR : {what}")

    lib.scotchfstratinit_(P(strat), ctypes.byref(ier)); check("Cannot initialize strategy")
    lib.scotchfrandomreset_()
    lib.scotchfgraphinit_(P(graph), ctypes.byref(ier)); check("Cannot initialize graph")
    lib.scotchfgraphbuild_(P(graph), I(0), I(nspec), P(xadj), P(xadj), P(load), P(xadj), I(int(xadj[nspec])), P(adjncy), P(adjncy),
                           ctypes.byref(ier)); check("Cannot build graph")
    lib.scotchfgraphcheck_(P(graph), ctypes.byref(ier)); check("Invalid check")
    lib.scotchfgraphpart_(P(graph), I(nparts), P(strat), P(part), ctypes.byref(ier)); check("Cannot part graph")
    lib.scotchfgraphexit_(P(graph), ctypes.byref(ier))
    lib.scotchfstratexit_(P(strat), ctypes.byref(ier))
    if part.min() < 0 or part.max() >= nparts:
        raise RuntimeError("Scotch returned an invalid partition number")
    return part
