"""Seismogram files the way the reference solver writes them, so that a run of this
harness can be compared with a REF_SEIS directory by the reference's own tooling
(utils/scripts/compare_seismogram_correlations.py reads exactly these two-column ASCII files).

Restated:
* file name  ``NET.STA.<band><comp>.sem{d,v,a,p}``          src/specfem3D/write_seismograms.f90:1013-1030
* band / instrument code from the sampling interval          src/specfem3D/write_seismograms.f90:1132-1160
* channel letter X/Y/Z (Cartesian, SUPPRESS_UTM_PROJECTION) or P   :1093-1107
* one line per sample ``time value`` with time = (it-1)*DT*NTSTEP_BETWEEN_OUTPUT_SAMPLE - t0, list-directed real output
                                                             src/specfem3D/write_output_ASCII_or_binary.f90:100-140
"""
from __future__ import annotations

import os
from typing import Dict, List

import numpy as np


def band_instrument_code(DT: float) -> str:
    if DT >= 1.0:
        return "LX"
    if DT > 0.1:
        return "MX"
    if DT > 0.0125:
        return "BX"
    if DT > 0.004:
        return "HX"
    if DT > 0.001:
        return "CX"
    return "FX"


def _fortran_real(v: float) -> str:
    """A default real the way gfortran's list-directed output shows it: 9 significant digits, plain notation for
    0.1 <= |v| < 1e9 (e.g. -9.94999981), exponent notation otherwise (e.g. 1.00000002E-24)."""
    v = float(np.float32(v))
    a = abs(v)
    if a == 0.0:
        return "0.00000000"
    if 0.1 <= a < 1.0e9:
        digits = 8 - int(np.floor(np.log10(a))) if a >= 1.0 else 9
        return f"{v:.{max(digits, 0)}f}"
    return f"{v:.8E}"


def write_seismograms(sim, seis: np.ndarray, out_dir: str, which: str = "d") -> List[str]:
    """``seis``: (nrec, nlength, 3) for d / v / a, or (nrec, nlength) for pressure (which = 'p').  Returns the paths."""
    os.makedirs(out_dir, exist_ok=True)
    rec = sim.receivers
    t = sim.time_axis()
    bic = band_instrument_code(sim.DT)  # sampling_rate = DT only (write_seismograms.f90:1077-1078)
    paths = []
    for irec in range(rec.nrec):
        comps = "P" if which == "p" else "XYZ"
        for c, ch in enumerate(comps):
            tr = seis[irec] if which == "p" else seis[irec, :, c]
            st = rec.stations[irec]
            path = os.path.join(out_dir, f"{st.network}.{st.name}.{bic}{ch}.sem{which}")
            with open(path, "w") as f:
                for ti, vi in zip(t, tr):
                    f.write(f"  {_fortran_real(ti):>12}     {_fortran_real(vi):>16}\n")
            paths.append(path)
    return paths


def read_seismogram(path: str) -> np.ndarray:
    """(nsamples, 2) time / value, the reader every REF_SEIS consumer uses (numpy.loadtxt in the reference scripts)."""
    return np.loadtxt(path, ndmin=2)


def compare_with_reference(syn_dir: str, ref_dir: str) -> Dict[str, Dict[str, float]]:
    """The three numbers compare_seismogram_correlations.py prints per trace (utils/scripts/compare_seismogram_correlations.py
    :120-200): zero-lag cross-correlation coefficient, relative L2 error and the time shift of the correlation maximum in
    samples.  CI accepts corr >= 0.999."""
    out = {}
    for name in sorted(os.listdir(ref_dir)):
        if ".sem" not in name or not os.path.exists(os.path.join(syn_dir, name)):
            continue
        r = read_seismogram(os.path.join(ref_dir, name))[:, 1]
        s = read_seismogram(os.path.join(syn_dir, name))[:, 1]
        n = min(len(r), len(s))
        r, s = r[:n], s[:n]
        nr, ns = np.linalg.norm(r), np.linalg.norm(s)
        corr = float(np.dot(r, s) / (nr * ns)) if nr > 0 and ns > 0 else 1.0
        err = float(np.linalg.norm(s - r) / nr) if nr > 0 else 0.0
        cc = np.correlate(s, r, mode="full") if n <= 20000 else np.array([np.dot(r, s)])
        shift = int(np.argmax(cc) - (n - 1)) if len(cc) > 1 else 0
        out[name] = {"corr": corr, "rel_l2": err, "shift": shift}
    return out


# ------------------------------------------------------------------------------------------------------------------
# adjoint sources of a kernel run
# ------------------------------------------------------------------------------------------------------------------
def write_adjoint_sources(sim, adj: np.ndarray, sem_dir: str) -> List[str]:
    """``SEM/NET.STA.<band>{X,Y,Z}.adj``: two-column ASCII ``time value`` like the seismograms, NSTEP lines per file."""
    os.makedirs(sem_dir, exist_ok=True)
    rec = sim.receivers
    # adjoint sources have NSTEP samples whatever NTSTEP_BETWEEN_OUTPUT_SAMPLE is
    t0 = sim.sources.t0 if sim.sources is not None else 0.0
    t = np.arange(sim.NSTEP) * sim.DT - t0
    bic = band_instrument_code(sim.DT)  # sampling_rate = DT only (write_seismograms.f90:1077-1078)
    paths = []
    for irec in range(rec.nrec):
        st = rec.stations[irec]
        for c, ch in enumerate("XYZ"):
            path = os.path.join(sem_dir, f"{st.network}.{st.name}.{bic}{ch}.adj")
            with open(path, "w") as f:
                for ti, vi in zip(t, adj[irec, :, c]):
                    f.write(f"  {_fortran_real(ti):>12}     {_fortran_real(vi):>16}\n")
            paths.append(path)
    return paths


def read_adjoint_sources(sim, sem_dir: str) -> np.ndarray:
    """compute_arrays_adjoint_source (src/specfem3D/compute_arrays_source.f90:203-291) for one chunk
    (NTSTEP_BETWEEN_READ_ADJSRC = NSTEP): the second column of the three component files of every receiver, NSTEP samples
    each (shorter files are an error, like the reference's 'file too short' stop), as (nrec, NSTEP, 3) float32.  The
    reference stores them as read, in mesh x / y / z (SUPPRESS_UTM_PROJECTION) -- no rotation is applied here either."""
    rec = sim.receivers
    bic = band_instrument_code(sim.DT)  # sampling_rate = DT only (write_seismograms.f90:1077-1078)
    out = np.zeros((rec.nrec, sim.NSTEP, 3), np.float32)
    for irec in range(rec.nrec):
        st = rec.stations[irec]
        for c, ch in enumerate("XYZ"):
            path = os.path.join(sem_dir, f"{st.network}.{st.name}.{bic}{ch}.adj")
            a = read_seismogram(path)
            if a.shape[0] < sim.NSTEP:
                raise ValueError(f"{path}: file too short ({a.shape[0]} lines), NSTEP = {sim.NSTEP}")
            out[irec, :, c] = a[: sim.NSTEP, 1].astype(np.float32)
    return out
