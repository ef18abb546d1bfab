"""The native host loop for SAVE_FORWARD runs and kernel runs (csrc/host_loop.cu: specfem_b200_iterate_time with SAVE_FORWARD,
specfem_b200_iterate_time_kernel_run, specfem_b200_iterate_time_undoatt) on the device against the Python loops.

These two paths were written after the round's GPU minutes were spent: on the CPU they issue exactly the calls, arguments and
records of the Python loops (tests/test_host_loops_call_order.py, mock of the ABI), but they have not run on a GPU yet.  Hence
(i) xfail(strict=False): a pass shows up as XPASS, a failure does not turn the suite red; (ii) the comparison runs in a CHILD
process (this file as a script), because the library reports errors with exit(1) like the reference does, which must not take the
test session down; (iii) the file name sorts last."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="first run on a GPU; CPU-verified against the Python loops on a mock ABI")
@pytest.mark.parametrize("nproc", ["1x1", "2x2"])
def test_native_forward_and_kernel_run_equal_the_python_loops(device_lib, nproc):
    res = subprocess.run([sys.executable, os.path.abspath(__file__), nproc], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True, timeout=600)
    print(res.stdout[-2000:])
    assert res.returncode == 0 and "NATIVE RUNS EQUAL" in res.stdout, res.stdout[-3000:]


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="first run on a GPU; CPU-verified against the Python loop on a mock ABI")
@pytest.mark.parametrize("case", ["undoatt_elastic_2x2", "undoatt_visco_2x1"])
def test_native_undo_attenuation_run_equals_the_python_loop(device_lib, case):
    res = subprocess.run([sys.executable, os.path.abspath(__file__), case], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True, timeout=600)
    print(res.stdout[-2000:])
    assert res.returncode == 0 and "NATIVE RUNS EQUAL" in res.stdout, res.stdout[-3000:]


def _child_undoatt(case):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref_cuda_runner as R
    from specfem3d_b200 import abi
    from specfem3d_b200.undo_att import UndoAttKernelRun, run_forward_and_save_undoatt
    from test_oracle_kernel_run import adjoint_source, rel_l2

    lib = abi.load()
    visco, nproc = case.endswith("visco_2x1"), ((2, 1) if case.endswith("2x1") else (2, 2))
    nstep = 120
    fwd, prod = run_forward_and_save_undoatt(R.undoatt_kernel_sim(nstep, 1, visco, nproc), lib=lib)
    adj = adjoint_source(fwd.gather_seismograms("d"))
    fwd.cleanup()
    out = {}
    for native in (False, True):
        k = UndoAttKernelRun(R.undoatt_kernel_sim(nstep, 3, visco, nproc), prod, adj, lib=lib)
        k.run_native() if native else k.run()
        out[native] = (k.kernels(), k.gather_seismograms("d"))
        k.cleanup()
    a, b = out[False], out[True]
    for ka, kb in zip(a[0], b[0]):
        for x, y in zip(ka, kb):
            assert np.abs(x).max() > 0 and rel_l2(y, x) <= 1e-4, "kernels"
    assert rel_l2(b[1], a[1]) <= 1e-4, "seismograms of the re-computed wavefield"
    print("NATIVE RUNS EQUAL: kernels", [rel_l2(y, x) for x, y in zip(a[0][0], b[0][0])])


def _child(nproc):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from specfem3d_b200 import abi
    from specfem3d_b200.kernel_run import KernelRun, run_forward_and_save
    from test_oracle_kernel_run import adjoint_source, kernel_sim, rel_l2

    lib = abi.load()
    nstep = 120
    out = {}
    for native in (False, True):
        fwd, prod = run_forward_and_save(kernel_sim(nproc, nstep), lib=lib, native=native)
        seis = fwd.gather_seismograms("d")
        fwd.cleanup()
        if not native:
            adj = adjoint_source(seis)
        k = KernelRun(kernel_sim(nproc, nstep, sim_type=3), prod, adj, lib=lib)
        if native:
            k.run_native(70)
            k.run_native()
        else:
            k.run()
        out[native] = (seis, [np.array(a) for a in prod.absorb], k.kernels(), k.gather_seismograms("d"))
        k.cleanup()
    a, b = out[False], out[True]
    assert np.abs(a[0]).max() > 0 and rel_l2(b[0], a[0]) <= 2e-5, "forward seismograms"
    for x, y in zip(a[1], b[1]):  # the Stacey records of every step
        assert x.shape == y.shape and x.shape[0] == nstep and rel_l2(y, x) <= 1e-4, "Stacey records"
    for ka, kb in zip(a[2], b[2]):  # rho / mu / kappa kernels of every rank
        for x, y in zip(ka, kb):
            assert np.abs(x).max() > 0 and rel_l2(y, x) <= 1e-4, "kernels"
    assert rel_l2(b[3], a[3]) <= 1e-4, "seismograms of the reconstructed wavefield"
    print("NATIVE RUNS EQUAL: forward seismograms", rel_l2(b[0], a[0]), "kernels", [rel_l2(y, x) for x, y in zip(a[2][0], b[2][0])])


if __name__ == "__main__":
    if sys.argv[1].startswith("undoatt"):
        _child_undoatt(sys.argv[1])
    else:
        px, py = (int(v) for v in sys.argv[1].split("x"))
        _child((px, py))
