"""The native host loop (specfem3d_b200/csrc/host_loop.cu) without a GPU: compiled with g++ against a MOCK of the C ABI that
only records which entry point was called for which rank (and `iphase`), and checked against the call order of the Fortran solver
it restates (iterate_time.F90:219-330; compute_forces_acoustic_calling_routine.f90:28-330; compute_forces_viscoelastic_calling_routine.
F90:30-435; assemble_MPI_vector_send_cuda / _write_cuda, assemble_MPI_vector.f90:640-830; write_seismograms; check_stability)."""
import ctypes
import json
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    d = tmp_path_factory.mktemp("host_loop")
    table = json.load(open(os.path.join(ROOT, "specfem3d_b200", "abi_symbols.json")))
    src = open(os.path.join(ROOT, "specfem3d_b200", "csrc", "host_loop.cu")).read()
    used = sorted(set(re.findall(r"\b([a-z_0-9]+)_\(", src)) & set(table))
    mock = ['#include "specfem3d_b200.h"', "#include <cstdio>", "#include <cstring>", "static char g_log[1 << 20]; static size_t g_n = 0;",
            'static void rec(const char* name, long* h, int iphase) { g_n += snprintf(g_log + g_n, sizeof(g_log) - g_n, "%s:%ld:%d\\n", name, h ? *h : -1L, iphase); }',
            'extern "C" const char* mock_log() { return g_log; }', 'extern "C" void mock_reset() { g_n = 0; g_log[0] = 0; }']
    for name in used:
        args = table[name]["args"]
        names = [a.replace("*", " ").split()[-1] for a in args]
        handle = next(n for n in names if n.startswith("Mesh_pointer"))
        iphase = next(("*" + n for n in names if n.lower().startswith("iphase")), "0")
        body = f'rec("{name}", {handle}, {iphase});'
        if name.startswith("get_norm"):
            body += " *norm = 1.5f + (float)*" + handle + ";"
        if name == "sync_copy_from_device":  # what the device would hand back: a value that identifies the sending rank
            body += " send_buffer[0] = 100.f + (float)*" + handle + ";"
        mock.append(f'extern "C" void {name}_({", ".join(args)}) {{ {body} }}')
    (d / "mock.cpp").write_text("\n".join(mock) + "\n")
    so = d / "libhost_mock.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-x", "c++",
                    os.path.join(ROOT, "specfem3d_b200", "csrc", "host_loop.cu"), str(d / "mock.cpp"), "-o", str(so)], check=True)
    return ctypes.CDLL(str(so))


def _run(host, nranks, elastic, acoustic, **flags):
    from specfem3d_b200.time_loop import _HostRank, _HostRun, _pf, _pi

    ranks = (_HostRank * nranks)()
    keep = []
    for r in range(nranks):
        h = ctypes.c_long(r)  # the "handle" is simply the rank number
        nb = np.array([1 - r], np.int32)
        nib = np.array([3], np.int32)
        ibo = np.array([[1, 2, 3]], np.int32)
        bufs = [np.zeros((1, 3, 3), np.float32) for _ in range(2)] + [np.zeros((1, 3), np.float32) for _ in range(2)]
        seis = [np.zeros((4, 1, 3), np.float32) for _ in range(4)]
        absb = np.zeros(4, np.float32)
        keep.append((h, nb, nib, ibo, bufs, seis, absb))
        k = ranks[r]
        k.Mesh_pointer, k.myrank, k.elastic, k.acoustic = ctypes.pointer(h), r, int(elastic), int(acoustic)
        k.nspec_outer_elastic, k.nspec_inner_elastic, k.nspec_outer_acoustic, k.nspec_inner_acoustic = 2, 5, 1, 3
        k.num_coupling_ac_el_faces = 4 if (elastic and acoustic) else 0
        k.num_interfaces_ext_mesh, k.max_nibool_interfaces_ext_mesh = (1, 3) if nranks > 1 else (0, 1)
        k.nibool_interfaces_ext_mesh, k.ibool_interfaces_ext_mesh, k.my_neighbors_ext_mesh = _pi(nib), _pi(ibo), _pi(nb)
        k.buffer_send_vector, k.buffer_recv_vector, k.buffer_send_scalar, k.buffer_recv_scalar = (_pf(b) for b in bufs)
        k.b_absorb_field = k.b_absorb_potential = _pf(absb)
        k.nrec_local = 1 if r == 0 else 0
        k.seismograms_d, k.seismograms_v, k.seismograms_a, k.seismograms_p = (_pf(s) for s in seis)
        k.seismo_current = 0
    run = _HostRun()
    stf = np.zeros((4, 1))
    norms = np.zeros(4, np.float32)
    run.nranks, run.ranks, run.NPROC = nranks, ranks, nranks
    run.deltat, run.deltatover2, run.deltatsqover2, run.NSTEP = 0.01, 0.005, 0.00005, 4
    run.ELASTIC_SIMULATION, run.ACOUSTIC_SIMULATION = int(elastic), int(acoustic)
    run.STACEY_ABSORBING_CONDITIONS = 1
    run.NSOURCES, run.stf_elastic, run.stf_acoustic = 1, stf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), stf.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    run.has_receivers, run.NTSTEP_BETWEEN_OUTPUT_SAMPLE, run.nlength_seismogram = 1, 1, 4
    run.host_exchange = 1 if nranks > 1 else 0
    run.norm_every, run.norms = 1, _pf(norms)
    for k, v in flags.items():
        setattr(run, k, v)
    host.mock_reset()
    host.specfem_b200_iterate_time_(ctypes.byref(run), ctypes.byref(ctypes.c_int(1)), ctypes.byref(ctypes.c_int(1)))
    host.mock_log.restype = ctypes.c_char_p
    log = [tuple(x.split(":")) for x in host.mock_log().decode().split()]
    return [(n, int(h), int(p)) for n, h, p in log], keep, norms, [ranks[r].seismo_current for r in range(nranks)]


def test_elastic_step_on_two_ranks_follows_the_fortran_order(host):
    log, keep, norms, seismo = _run(host, 2, True, False)
    F, S, SRC, PACK = "compute_forces_viscoelastic_cuda", "compute_stacey_viscoelastic_cuda", "compute_add_sources_el_cuda", "transfer_boundary_from_device_a"
    assert log == [
        ("update_displacement_cuda", 0, 0), ("update_displacement_cuda", 1, 0),
        # iphase 1: outer elements, Stacey, sources, then the interface values leave the device (every rank in turn)
        (F, 0, 1), (S, 0, 1), (SRC, 0, 0), (PACK, 0, 0), (F, 1, 1), (S, 1, 1), (SRC, 1, 0), (PACK, 1, 0),
        # iphase 2: inner elements while the messages travel, then receive and assemble
        (F, 0, 2), (F, 1, 2), ("sync_copy_from_device", 0, 2), ("sync_copy_from_device", 1, 2),
        ("transfer_boundary_to_device_a", 0, 0), ("transfer_boundary_to_device_a", 1, 0),
        ("transfer_asmbl_accel_to_device", 0, 0), ("transfer_asmbl_accel_to_device", 1, 0),
        ("kernel_3_a_cuda", 0, 0), ("kernel_3_a_cuda", 1, 0),
        ("compute_seismograms_cuda", 0, 0),  # only the rank that holds receivers
        ("get_norm_elastic_from_device", 0, 0), ("get_norm_elastic_from_device", 1, 0)]
    # the "MPI" between the two ranks of the process: each one's send buffer arrived in the other's receive buffer
    assert keep[0][4][1][0, 0, 0] == 101.0 and keep[1][4][1][0, 0, 0] == 100.0
    assert norms[0] == 2.5 and seismo == [1, 1]  # max over the ranks; every rank counts the sample


def test_coupled_step_on_one_rank_follows_the_fortran_order(host):
    log, _, _, _ = _run(host, 1, True, True, norm_every=0)
    names = [(n, p) for n, _, p in log]
    assert names == [
        ("update_displacement_ac_cuda", 0), ("update_displacement_cuda", 0),  # update_displ_Newmark: acoustic, then elastic
        # acoustic solve first (compute_forces_acoustic_calling): its coupling term needs the displacement just predicted
        ("acoustic_enforce_free_surf_cuda", 0), ("compute_forces_acoustic_cuda", 1), ("compute_stacey_acoustic_cuda", 1),
        ("compute_coupling_ac_el_cuda", 1), ("compute_add_sources_ac_cuda", 0), ("compute_forces_acoustic_cuda", 2),
        ("kernel_3_acoustic_cuda", 0), ("acoustic_enforce_free_surf_cuda", 0),
        # elastic solve (compute_forces_viscoelastic_calling): coupled to the pressure just computed
        ("compute_forces_viscoelastic_cuda", 1), ("compute_stacey_viscoelastic_cuda", 1), ("compute_coupling_el_ac_cuda", 1),
        ("compute_add_sources_el_cuda", 0), ("compute_forces_viscoelastic_cuda", 2), ("kernel_3_a_cuda", 0),
        ("compute_seismograms_cuda", 0)]


def test_ocean_load_and_cpml_branches(host):
    log, _, _, _ = _run(host, 1, True, False, norm_every=0, APPROXIMATE_OCEAN_LOAD=1, PML_CONDITIONS=1, STACEY_ABSORBING_CONDITIONS=0)
    names = [n for n, _, _ in log]
    assert names == ["update_displacement_cuda", "compute_forces_viscoelastic_cuda", "compute_add_sources_el_cuda",
                     "compute_forces_viscoelastic_cuda", "pml_impose_boundary_condition_elastic_cuda", "kernel_3_a_cuda",
                     "compute_coupling_ocean_cuda", "kernel_3_b_cuda", "compute_seismograms_cuda"]
