"""The two callers of the C ABI -- the Python time loop (specfem3d_b200/time_loop.py, the one the oracle tests pin on the GPU) and the
native host loop (csrc/host_loop.cu) -- must issue the SAME entry points for the same ranks in the same order.  Checked without a
GPU: both drive a MOCK library that exports all 124 entry points of the contract (generated from abi_symbols.json; every call is
recorded as (name, handle, iphase, weighted sum of the first elements of its scalar-typed arguments)) together with the compiled host loop.  tests/test_gpu_parity.py::
test_native_host_loop_equals_python_loop compares the numbers on the device; this compares the call sequences."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mock_lib(tmp_path_factory):
    from specfem3d_b200 import abi

    d = tmp_path_factory.mktemp("mock_abi")
    table = json.load(open(os.path.join(ROOT, "specfem3d_b200", "abi_symbols.json")))
    src = ['#include "specfem3d_b200.h"', "#include <cstdio>", "static char g_log[1 << 22]; static size_t g_n = 0; static long g_handles = 0; static long g_records = 0;",
           'static void rec(const char* name, long* h, int iphase, double args) { if (g_n + 160 < sizeof(g_log)) g_n += snprintf(g_log + g_n, 160, "%s:%ld:%d:%.9e\\n", name, h ? *h : -1L, iphase, args); }',
           'extern "C" const char* mock_log() { return g_log; }', 'extern "C" void mock_reset(int handles_too) { g_n = 0; g_log[0] = 0; if (handles_too) g_handles = 0; g_records = 0; }']
    for name, entry in table.items():
        args = entry["args"]
        names = [a.replace("*", " ").replace("[]", " ").split()[-1] for a in args]
        handle = next((n for n in names if n.startswith("Mesh_pointer")), None)
        iphase = next(("*" + n for n in names if n.lower().startswith("iphase")), "0")
        body = ""
        if name == "prepare_constants_device":
            body += f"*{handle} = ++g_handles; "
        # scalar arguments arrive by reference: the first element of every int / realw / double pointer goes into the record too
        # (time-step constants, element counts, FORWARD_OR_ADJOINT, the source time function value of the step, ...); in the
        # time loop every such pointer refers to at least one element
        first = []
        if name not in ("initialize_gpu_device", "prepare_constants_device") and not name.startswith(("prepare_", "get_")):
            for a, n in zip(args, names):
                t = a.replace("const", " ").replace("*", " ").split()[0]
                if "*" in a and t in ("int", "realw", "double", "field") and not n.startswith("Mesh_pointer"):
                    first.append(f"({n} ? (double)*{n} : -12345.0)")
        body += f'rec("{name}", {handle or "nullptr"}, {iphase}, {" + ".join(f"{k + 1}.0 * {e}" for k, e in enumerate(first)) or "0.0"});'
        if name.startswith("get_norm"):
            body += " *norm = 1.0f;"
        if name in ("compute_stacey_viscoelastic_cuda", "compute_stacey_acoustic_cuda"):  # the record the device would hand back
            rec_arg = next(n for n in names if "b_absorb" in n)
            body += f" {rec_arg}[0] = (float)(++g_records) + 0.25f * (float)*{handle};"
        if name == "initialize_gpu_device":
            body += " *ncuda_devices = 1;"
        src.append(f'extern "C" void {name}_({", ".join(args) if args else "void"}) {{ {body} }}')
    (d / "mock.cpp").write_text("\n".join(src) + "\n")
    so = d / "libmock_abi.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-x", "c++",
                    os.path.join(ROOT, "specfem3d_b200", "csrc", "host_loop.cu"), str(d / "mock.cpp"), "-o", str(so)], check=True)
    lib = abi.DeviceLibrary(str(so))
    lib.cdll.mock_log.restype = ctypes.c_char_p
    return lib


def _calls(lib):
    return [tuple(x.split(":")) for x in lib.cdll.mock_log().decode().split()]


def _cases():
    import test_gpu_parity as g  # the builders of the GPU tests (importing the module needs no GPU)

    return {"elastic": lambda: g.small_sim(nstep=6, nex=(4, 4, 2)),
            "elastic_2x2": lambda: g.small_sim(nproc=(2, 2), nstep=6, nex=(4, 4, 2)),
            "viscoelastic": lambda: g.visco_sim(nstep=6),
            "coupled_2x1": lambda: g.ocean_sim(nproc=(2, 1), nstep=6),
            "cpml": lambda: g.pml_sim(nstep=6),
            "cpml_2x2": lambda: g.pml_sim(nproc=(2, 2), nstep=6),
            "ocean_load": lambda: g.ocean_load_sim(nstep=6),
            "ocean_load_2x2": lambda: g.ocean_load_sim(nproc=(2, 2), nstep=6),
            "coloring": lambda: g.small_sim(nstep=6, nex=(4, 4, 2), coloring=True),
            "coupled_coloring": lambda: g.ocean_sim(nstep=6, coloring=True),
            "anisotropic_visco": lambda: g.aniso_sim(nstep=6, attenuation=True, nproc=(2, 1)),
            "source_in_water": lambda: g.ocean_sim(nproc=(2, 1), nstep=6, source_in_water=True)}


@pytest.mark.parametrize("case", ["elastic", "elastic_2x2", "viscoelastic", "coupled_2x1", "cpml", "cpml_2x2", "ocean_load", "ocean_load_2x2",
                                  "coloring", "coupled_coloring", "anisotropic_visco", "source_in_water"])
def test_python_and_native_host_loops_issue_the_same_calls(mock_lib, case):
    from specfem3d_b200.time_loop import DeviceRun

    logs = {}
    for native in (False, True):
        sim = _cases()[case]()
        mock_lib.cdll.mock_reset(1)
        run = DeviceRun(sim, lib=mock_lib)
        prepared = _calls(mock_lib)
        assert prepared[0][0] == "initialize_gpu_device" and any(c[0] == "prepare_constants_device" for c in prepared)
        mock_lib.cdll.mock_reset(0)
        if native:
            run.run_native(4)
            run.run_native(2)  # resumes where it stopped
        else:
            run.run(6)
        logs[native] = _calls(mock_lib)
        run.cleanup()
    a, b = logs[False], logs[True]
    assert len(a) > 6 * 4 and any(c[0].startswith("compute_forces") for c in a)
    assert a == b, next((i, x, y) for i, (x, y) in enumerate(zip(a, b)) if x != y)


@pytest.mark.parametrize("case", ["elastic_2x2", "coupled_2x1"])
def test_save_forward_records_are_kept_alike(mock_lib, case):
    """SIMULATION_TYPE 1 with SAVE_FORWARD: both loops keep the Stacey record every compute_stacey_* call hands back
    (write_abs(IOABS, b_absorb_field, b_reclen_field, it)) -- the mock numbers the records, so equal histories mean: same calls in
    the same order, each record stored under its own step."""
    from specfem3d_b200.time_loop import DeviceRun

    hist = {}
    for native in (False, True):
        sim = _cases()[case]()
        sim.SAVE_FORWARD = True
        mock_lib.cdll.mock_reset(1)
        run = DeviceRun(sim, lib=mock_lib)
        mock_lib.cdll.mock_reset(0)
        if native:
            run.run_native(4)
            run.run_native(2)
        else:
            run.run(6)
        hist[native] = ([np.array(r.absorb_history) for r in run.ranks], [np.array(r.absorb_potential_history) for r in run.ranks],
                        _calls(mock_lib))
        run.cleanup()
    a, b = hist[False], hist[True]
    assert a[2] == b[2]
    n_records = 0
    for x, y in zip(a[0] + a[1], b[0] + b[1]):
        assert x.shape == y.shape and np.array_equal(x, y)
        n_records += len(x)
        if len(x):
            assert len(x) == 6 and len(np.unique(x.reshape(6, -1)[:, 0])) == 6  # six steps, six different records
    assert n_records >= 6 * len(a[0])


@pytest.mark.parametrize("case", ["elastic", "elastic_2x2", "acoustic_2x2", "coupled", "elastic_hess"])
def test_kernel_run_loops_issue_the_same_calls(mock_lib, case):
    """SIMULATION_TYPE 3: KernelRun.step (kernel_run.py, the loop the GPU tests pin against the oracle and the reference CUDA backend)
    and specfem_b200_iterate_time_kernel_run (csrc/host_loop.cu) -- same entry points, ranks, iphase and scalar arguments, incl. the
    Stacey record replayed at every step (read_abs at NSTEP - it + 1: the mock numbered the records of the forward run), the
    adjoint-source array, the source time function taken backwards, and the last frame handed over after the first step."""
    import test_oracle_kernel_run as k

    from specfem3d_b200.kernel_run import KernelRun, run_forward_and_save

    nstep = 6
    make = {"elastic": lambda t: k.kernel_sim((1, 1), nstep, nex=(4, 4, 2), sim_type=t),
            "elastic_2x2": lambda t: k.kernel_sim((2, 2), nstep, nex=(4, 4, 2), sim_type=t),
            "elastic_hess": lambda t: k.kernel_sim((1, 1), nstep, nex=(4, 4, 2), sim_type=t),
            "acoustic_2x2": lambda t: k.acoustic_kernel_sim((2, 2), nstep, nex=(4, 4, 2), sim_type=t),
            "coupled": lambda t: k.coupled_kernel_sim(nstep, t)}[case]
    logs = {}
    for native in (False, True):
        mock_lib.cdll.mock_reset(1)
        fwd, prod = run_forward_and_save(make(1), lib=mock_lib, native=native)
        fwd.cleanup()
        for ir, last in enumerate(prod.last):  # a recognisable last frame (the mock's device hands back zeros)
            for j, key in enumerate(sorted(last)):
                last[key] = np.full_like(last[key], 10.0 * (ir + 1) + j)
        sim3 = make(3)
        sim3.APPROXIMATE_HESS_KL = case == "elastic_hess"
        rng = np.random.default_rng(5)
        adj = rng.standard_normal((sim3.receivers.nrec, nstep, 3)).astype(np.float32)
        mock_lib.cdll.mock_reset(1)
        run = KernelRun(sim3, prod, adj, lib=mock_lib)
        mock_lib.cdll.mock_reset(0)
        if native:
            run.run_native(4)
            run.run_native()
        else:
            run.run()
        assert run.it == nstep
        logs[native] = _calls(mock_lib)
        run.cleanup()
    a, b = logs[False], logs[True]
    names = {c[0] for c in a}
    assert {"compute_kernels_elastic_cuda", "compute_kernels_acoustic_cuda"} & names and any(n.startswith("transfer_b_fields") for n in names)
    assert a == b, next((i, x, y) for i, (x, y) in enumerate(zip(a, b)) if x != y)


@pytest.mark.parametrize("case", ["elastic", "elastic_2x2", "viscoelastic_2x1", "acoustic_2x2"])
def test_undo_attenuation_loops_issue_the_same_calls(mock_lib, case):
    """UNDO_ATTENUATION_AND_OR_PML kernel runs: UndoAttKernelRun.run (undo_att.py, iterate_time_undoatt restated; the loop the GPU
    tests pin) and specfem_b200_iterate_time_undoatt (csrc/host_loop.cu) -- same entry points, ranks, iphase and scalar arguments over
    the whole run: subsets of 4, 4 and 3 steps taken in reverse order, each re-computed forward and then swept by the adjoint wavefield
    with the buffered frames handed back in reverse order."""
    import ref_cuda_runner as R

    from specfem3d_b200.undo_att import UndoAttKernelRun, run_forward_and_save_undoatt

    nstep, nt_dump = 11, 4
    make = {"elastic": lambda t: R.undoatt_kernel_sim(nstep, t, False, (1, 1), nt_dump),
            "elastic_2x2": lambda t: R.undoatt_kernel_sim(nstep, t, False, (2, 2), nt_dump),
            "viscoelastic_2x1": lambda t: R.undoatt_kernel_sim(nstep, t, True, (2, 1), nt_dump),
            "acoustic_2x2": lambda t: R.undoatt_acoustic_sim(nstep, t, (2, 2), nt_dump)}[case]
    logs, fwd_logs = {}, {}
    for native in (False, True):
        mock_lib.cdll.mock_reset(1)
        fwd, prod = run_forward_and_save_undoatt(make(1), lib=mock_lib, native=native)  # subset by subset, frames saved in between
        forward_log = [c for c in _calls(mock_lib) if not c[0].startswith(("prepare_", "initialize_"))]
        fwd.cleanup()
        for ir, frames in enumerate(prod.frames):  # recognisable frames (the mock's device hands back zeros)
            for s, f in enumerate(frames):
                for j, key in enumerate(sorted(f)):
                    if key == "R":
                        f[key] = [np.full_like(a, 100.0 * (ir + 1) + 10.0 * s + m) for m, a in enumerate(f[key])]
                    else:
                        f[key] = np.full_like(f[key], 10.0 * (ir + 1) + s + 0.1 * j)
        sim3 = make(3)
        rng = np.random.default_rng(7)
        adj = rng.standard_normal((sim3.receivers.nrec, nstep, 3)).astype(np.float32)
        mock_lib.cdll.mock_reset(1)
        run = UndoAttKernelRun(sim3, prod, adj, lib=mock_lib)
        assert (run.nsubset, run.lengths) == (3, [4, 4, 3])
        mock_lib.cdll.mock_reset(0)
        if native:
            run.run_native()
        else:
            run.run()
        assert run.it == nstep
        logs[native] = _calls(mock_lib)
        fwd_logs[native] = forward_log
        run.cleanup()
    a, b = logs[False], logs[True]
    fa, fb = fwd_logs[False], fwd_logs[True]
    assert any(c[0].endswith("_undoatt_cuda") for c in fa) and not any(c[0] in ("compute_stacey_viscoelastic_cuda", "compute_stacey_acoustic_cuda") for c in fa)
    assert fa == fb, next((i, x, y) for i, (x, y) in enumerate(zip(fa, fb)) if x != y)
    names = [c[0] for c in a]
    assert names.count("transfer_b_fields_to_device") + names.count("transfer_b_fields_ac_to_device") == 3 * len({c[1] for c in a})
    assert a == b, next((i, x, y) for i, (x, y) in enumerate(zip(a, b)) if x != y)
