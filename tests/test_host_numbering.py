"""Host-side C++ of the product that needs no GPU: the library's own point numbering (specfem3d_b200/csrc/renumber.h), compiled
with g++ into a scratch shared object and checked on harness databases: it is a bijection, the points of the elastic elements form
the prefix the domain-restricted Newmark passes rely on, and inside a group of five xi-neighbours a row of a
k-plane is one run of consecutive new ids (what the row kernels' coalesced gathers rely on)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from specfem3d_b200.configs import HALFSPACE
from specfem3d_b200.gll import GLL
from specfem3d_b200.meshfem import BoxSpec, Material, build_partition

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def numbering(tmp_path_factory):
    d = tmp_path_factory.mktemp("renumber")
    src = d / "shim.cpp"
    src.write_text('#include "renumber.h"\n#include <cstring>\n'
                   'extern "C" void point_numbering(const int* ibool, int nspec, int nglob, const int* inner, const int* el,\n'
                   '                                const int* ac, const int* irr, int* perm) {\n'
                   '  std::vector<int> p = sb200::compute_point_numbering(ibool, nspec, nglob, inner, el, ac, irr);\n'
                   '  std::memcpy(perm, p.data(), sizeof(int) * (size_t)nglob);\n}\n')
    so = d / "libshim.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "specfem3d_b200", "csrc"), "-o", str(so), str(src)],
                   check=True)
    lib = ctypes.CDLL(str(so))

    def run(db):
        ibool = np.ascontiguousarray(np.asarray(db.ibool).reshape(db.NSPEC_AB, 125), np.int32)
        inner = np.ascontiguousarray(db.ispec_is_inner, np.int32)
        el = np.ascontiguousarray(db.ispec_is_elastic, np.int32)
        ac = np.ascontiguousarray(db.ispec_is_acoustic, np.int32)
        irr = np.ascontiguousarray(db.irregular_element_number, np.int32)
        perm = np.full(db.NGLOB_AB, -7, np.int32)
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        lib.point_numbering(P(ibool), db.NSPEC_AB, db.NGLOB_AB, P(inner), P(el), P(ac), P(irr), P(perm))
        return ibool - 1, perm

    return run


def test_numbering_is_a_bijection_with_an_elastic_prefix(numbering):
    water = Material(1020.0, 1500.0, 0.0, 9999.0, 9999.0, 1)
    spec = BoxSpec(8, 6, 6, nproc_xi=2, nproc_eta=1, xmax=8800.0, ymax=5400.0, zmin=-6000.0, materials=(HALFSPACE, water),
                   layer_material=[0] * 4 + [1] * 2)
    gll = GLL()
    for rank in range(2):
        db = build_partition(spec, rank, gll)
        ibool, perm = numbering(db)
        assert np.array_equal(np.sort(perm), np.arange(db.NGLOB_AB))  # bijection on [0, NGLOB)
        el_pts = np.unique(perm[ibool[np.asarray(db.ispec_is_elastic) != 0]])
        ac_only = np.setdiff1d(np.unique(perm[ibool[np.asarray(db.ispec_is_acoustic) != 0]]), el_pts)
        assert np.array_equal(el_pts, np.arange(len(el_pts)))  # the solid's points are exactly the prefix [0, el_hi)
        assert ac_only.min() == len(el_pts) and np.array_equal(ac_only, np.arange(len(el_pts), db.NGLOB_AB))


def test_rows_of_a_group_are_runs_of_consecutive_points(numbering):
    spec = BoxSpec(10, 4, 4, xmax=11000.0, ymax=3600.0, zmin=-4000.0, materials=(HALFSPACE,))  # irregular, one rank: all inner
    db = build_partition(spec, 0, GLL())
    ibool, perm = numbering(db)
    new = perm[ibool].reshape(db.NSPEC_AB, 5, 5, 5)  # [element, k, j, i]
    # elements 0 ... 4 are five xi-neighbours (x fastest) and the first group of the launch order: row (k, j) through the five
    # elements holds 21 distinct points (shared faces counted once); wherever none of them was numbered before (j >= 1, k >= 1
    # skip the planes earlier rows / planes introduced) they are consecutive
    for k in range(5):
        for j in range(5):
            row = np.concatenate([new[e, k, j] for e in range(5)])
            ids = np.unique(row)
            assert len(ids) == 21
            if k >= 1 and j >= 1:
                assert ids.max() - ids.min() == 20
    assert new[:5].max() == len(np.unique(new[:5])) - 1  # the group's 21 * 25 points are the first 525 ids
