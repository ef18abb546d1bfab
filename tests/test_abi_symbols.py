"""The C-ABI shared library loads without a GPU and exports every symbol declared in include/specfem3d_b200.h
(i.e. all 124 entry points of the reference contract, gfortran-mangled, plus the documented extensions)."""
import ctypes
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from specfem3d_b200.build import build_cuda, LIB_PATH

    build_cuda()  # nvcc cross-compiles without a GPU
    return ctypes.CDLL(LIB_PATH)


def declared_symbols():
    names = []
    inc = os.path.join(ROOT, "include")
    for f in sorted(os.listdir(inc)):  # specfem3d_b200.h (the reference contract + extensions), specfem3d_b200_host.h (native host loop)
        if f.endswith(".h"):
            names += re.findall(r"void\s+SB200_FC\((\w+)\)\s*\(", open(os.path.join(inc, f)).read())
    return names


def test_header_declares_the_full_reference_contract():
    table = json.load(open(os.path.join(ROOT, "specfem3d_b200", "abi_symbols.json")))
    names = declared_symbols()
    assert len(table) == 124
    assert set(table) <= set(names)
    # the must-export set for forward runs is implemented, not stubbed
    for n in ("prepare_constants_device", "compute_forces_viscoelastic_cuda", "update_displacement_cuda", "kernel_3_a_cuda",
              "compute_stacey_viscoelastic_cuda", "compute_coupling_ac_el_cuda", "compute_coupling_el_ac_cuda",
              "transfer_boundary_from_device_a", "transfer_asmbl_accel_to_device", "compute_seismograms_cuda",
              "compute_forces_acoustic_cuda", "compute_add_sources_el_cuda", "get_norm_elastic_from_device"):
        assert table[n]["implemented"], n


def test_every_declared_symbol_is_exported(lib):
    missing = []
    for n in declared_symbols():
        try:
            getattr(lib, n + "_")
        except AttributeError:
            missing.append(n)
    assert not missing, missing


def test_fortran_style_binding_checks_arguments():
    from specfem3d_b200 import abi

    L = abi.load()
    with pytest.raises(TypeError):
        L.update_displacement_cuda(ctypes.c_long(0), 0.1)  # wrong arity is caught before touching the device


def _run_snippet(code):
    import subprocess
    import sys

    return subprocess.run([sys.executable, "-c", code], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)


def test_out_of_scope_entry_points_fail_loudly(tmp_path):
    """No silent no-ops and no CPU fallback: an entry point outside the implemented set prints its name and exits non-zero
    (like exit_on_error in src/gpu/helper_functions.cu), and so does an implemented one called with a null handle."""
    load = "import ctypes; L = ctypes.CDLL('specfem3d_b200/libspecfem3d_b200.so'); h = ctypes.c_long(0); i = ctypes.c_int(1)\n"
    res = _run_snippet(load + "L.get_max_accel_(ctypes.byref(i), ctypes.byref(i), ctypes.byref(h))\nprint('survived')")
    assert res.returncode != 0 and "survived" not in res.stdout
    assert "get_max_accel" in res.stderr + res.stdout
    f = "f = ctypes.c_float(0.1)\n"
    res = _run_snippet(load + f + "L.update_displacement_cuda_(ctypes.byref(h), ctypes.byref(f), ctypes.byref(f), ctypes.byref(f), ctypes.byref(i))\nprint('survived')")
    assert res.returncode != 0 and "survived" not in res.stdout
    assert "Mesh_pointer" in res.stderr + res.stdout
    table = json.load(open(os.path.join(ROOT, "specfem3d_b200", "abi_symbols.json")))
    # kernel runs, ocean load: implemented this round
    for n in ("prepare_fields_elastic_adj_dev", "compute_kernels_elastic_cuda", "compute_kernels_acoustic_cuda", "compute_coupling_ocean_cuda",
              "transfer_kernels_el_to_host", "add_sources_el_sim_type_2_or_3", "transfer_b_fields_to_device"):
        assert table[n]["implemented"], n


@pytest.mark.parametrize("compiler,std", [("gcc", "-std=c99"), ("g++", "-std=c++11")])
def test_headers_are_plain_c(tmp_path, compiler, std):
    """include/*.h is the boundary a C, C++ or Fortran caller binds to: plain pointers and sizes, no CUDA / torch types.  Both
    headers must compile on their own with a C and a C++ compiler, and a designated-initialiser use of the host-loop structs
     must type-check."""
    import subprocess

    src = tmp_path / ("t.c" if compiler == "gcc" else "t.cpp")
    src.write_text('#include "specfem3d_b200.h"\n#include "specfem3d_b200_host.h"\n'
                   "int main(void) {\n"
                   "  specfem_b200_host_rank rank;\n  specfem_b200_host_run run;\n"
                   "  long handle = 0; int one = 1;\n"
                   "  rank.Mesh_pointer = &handle; rank.myrank = 0;\n"
                   "  run.nranks = 1; run.ranks = &rank; run.USE_TRICK_FOR_BETTER_PRESSURE = 0;\n"
                   "  if (sizeof(realw) != 4) return 1;\n"
                   "  if (0) specfem_b200_iterate_time_(&run, &one, &one);\n"
                   "  return 0;\n}\n")
    res = subprocess.run([compiler, std, "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout[-3000:]
