"""Regenerates tests/golden/*.npz from the reference tree (run in the build container, where /root/reference exists).

halfspace_ref_seis.npz : the reference's own golden seismograms for the default example (config 1),
    EXAMPLES/applications/homogeneous_halfspace/REF_SEIS/DB.X{20,30,40,50}.BX{X,Y,Z}.semd, all 5000 samples
    (ASCII "time value" -> float32 arrays), plus the max-norm lines of REF_SEIS/output_solver.txt.
cpml_ref_seis.npz : the same for the reference's elastic C-PML example,
    EXAMPLES/applications/CPML_examples/homogeneous_halfspace_HEX8_elastic_absorbing_CPML_5sides/REF_SEIS/DB.X??.MX?.semd
    (all 2400 samples: the late part of the traces is where an absorbing layer shows).
"""
import os
import re
import sys

import numpy as np

REF = "/root/reference/EXAMPLES/applications/homogeneous_halfspace/REF_SEIS"
NKEEP = 5000  # all samples (round 1 kept 1500; the GPU test now runs the whole example)
HERE = os.path.dirname(os.path.abspath(__file__))


REF_CPML = "/root/reference/EXAMPLES/applications/CPML_examples/homogeneous_halfspace_HEX8_elastic_absorbing_CPML_5sides/REF_SEIS"


def maxnorms(path):
    steps, norms = [], []
    txt = open(path).read()
    for m in re.finditer(r"Time step #\s+(\d+).*?Max norm displacement vector U in all slices \(m\) =\s+([0-9.E+-]+)", txt, re.S):
        steps.append(int(m.group(1)))
        norms.append(float(m.group(2)))
    return np.array(steps, dtype=np.int32), np.array(norms, dtype=np.float64)


def cpml():
    out = {}
    for st in ("X20", "X30", "X40", "X50"):
        for ch in "XYZ":
            d = np.loadtxt(os.path.join(REF_CPML, f"DB.{st}.MX{ch}.semd"))
            out[f"{st}_{ch}"] = d[:, 1].astype(np.float32)
            out["time"] = d[:, 0].astype(np.float32)
    out["maxnorm_step"], out["maxnorm_value"] = maxnorms(os.path.join(REF_CPML, "output_solver.txt"))
    np.savez_compressed(os.path.join(HERE, "cpml_ref_seis.npz"), **out)
    print("wrote cpml_ref_seis.npz", len(out["time"]), "samples,", len(out["maxnorm_step"]), "max-norm lines")


REF_ATT = "/root/reference/EXAMPLES/benchmarks/attenuation/viscoelastic/REF_SEIS"


def attenuation():
    """attenuation_ref_seis.npz: the three traces of the viscoelastic benchmark (CE.1.FX{X,Y,Z}.semd, 1300 samples) and the
    max-norm lines of its output_solver.txt."""
    out = {}
    for ch in "XYZ":
        d = np.loadtxt(os.path.join(REF_ATT, f"CE.1.FX{ch}.semd"))
        out[f"1_{ch}"] = d[:, 1].astype(np.float32)
        out["time"] = d[:, 0].astype(np.float64)
    out["maxnorm_step"], out["maxnorm_value"] = maxnorms(os.path.join(REF_ATT, "output_solver.txt"))
    np.savez_compressed(os.path.join(HERE, "attenuation_ref_seis.npz"), **out)
    print("wrote attenuation_ref_seis.npz", len(out["time"]), "samples,", len(out["maxnorm_step"]), "max-norm lines")


REF_WATER = "/root/reference/EXAMPLES/applications/waterlayered_halfspace/REF_SEIS"


def waterlayer():
    """waterlayer_ref_seis.npz: the coupled fluid-solid example (EXAMPLES/applications/waterlayered_halfspace/REF_SEIS, only
    vertical components exist there): the nine displacement traces DB.X{10..90}.HXZ.semd, the velocity traces
    DB.*.HXZ.semv of all 22 stations (14 receivers in the water, 8 in the rock, 8 of them within 20 cm of the sea floor),
    4500 samples each, and the max-norm lines of output_solver.txt (elastic displacement, acoustic potential_dot_dot)."""
    out = {}
    for st in ("X10", "X20", "X30", "X40", "X50", "X60", "X70", "X80", "X90"):
        d = np.loadtxt(os.path.join(REF_WATER, f"DB.{st}.HXZ.semd"))
        out[f"d_{st}"] = d[:, 1].astype(np.float32)
        out["time"] = d[:, 0].astype(np.float64)
    for st in ("X10", "X20", "X30", "X40", "X50", "X55", "X60", "X70", "X80", "X90", "X1", "X2", "X3", "X4", "X5", "X6",
               "X7", "X8", "Y1", "Y2", "Y3", "Y4"):
        out[f"v_{st}"] = np.loadtxt(os.path.join(REF_WATER, f"DB.{st}.HXZ.semv"))[:, 1].astype(np.float32)
    txt = open(os.path.join(REF_WATER, "output_solver.txt")).read()
    steps, nu, na = [], [], []
    for m in re.finditer(r"Time step #\s+(\d+).*?Max norm displacement vector U in all slices \(m\) =\s+([0-9.E+-]+)\s*\n"
                         r"\s*Max norm pressure P in all slices \(Pa\) =\s+([0-9.E+-]+)", txt, re.S):
        steps.append(int(m.group(1)))
        nu.append(float(m.group(2)))
        na.append(float(m.group(3)))
    out["maxnorm_step"] = np.array(steps, dtype=np.int32)
    out["maxnorm_displ"] = np.array(nu, dtype=np.float64)
    out["maxnorm_pressure"] = np.array(na, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "waterlayer_ref_seis.npz"), **out)
    print("wrote waterlayer_ref_seis.npz", len(out["time"]), "samples,", len(steps), "max-norm lines")


REF_ACOUSTIC = "/root/reference/EXAMPLES/applications/homogeneous_acoustic/REF_SEIS"


def acoustic():
    """acoustic_ref_seis.npz: the purely acoustic example (EXAMPLES/applications/homogeneous_acoustic/REF_SEIS): the four
    PRESSURE traces DB.X{20,30,40,50}.MXP.semp (1000 samples, USE_TRICK_FOR_BETTER_PRESSURE) and the max-norm lines of
    output_solver.txt (potential_dot_dot)."""
    out = {}
    for st in ("X20", "X30", "X40", "X50"):
        d = np.loadtxt(os.path.join(REF_ACOUSTIC, f"DB.{st}.MXP.semp"))
        out[f"p_{st}"] = d[:, 1].astype(np.float32)
        out["time"] = d[:, 0].astype(np.float64)
    txt = open(os.path.join(REF_ACOUSTIC, "output_solver.txt")).read()
    steps, na = [], []
    for m in re.finditer(r"Time step #\s+(\d+).*?Max norm pressure P in all slices \(Pa\) =\s+([0-9.E+-]+)", txt, re.S):
        steps.append(int(m.group(1)))
        na.append(float(m.group(2)))
    out["maxnorm_step"] = np.array(steps, dtype=np.int32)
    out["maxnorm_pressure"] = np.array(na, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "acoustic_ref_seis.npz"), **out)
    print("wrote acoustic_ref_seis.npz", len(out["time"]), "samples,", len(steps), "max-norm lines")


REF_NOABS = "/root/reference/EXAMPLES/applications/homogeneous_halfspace_HEX8_elastic_no_absorbing/REF_SEIS"


def no_absorbing():
    """noabs_ref_seis.npz: config 1 without absorbing boundaries (EXAMPLES/applications/
    homogeneous_halfspace_HEX8_elastic_no_absorbing/REF_SEIS/DB.X{20,30,40,50}.BX{X,Y,Z}.semd, all 5000 samples) and the
    max-norm lines of its output_solver.txt."""
    out = {}
    for st in ("X20", "X30", "X40", "X50"):
        for ch in "XYZ":
            d = np.loadtxt(os.path.join(REF_NOABS, f"DB.{st}.BX{ch}.semd"))
            out[f"{st}_{ch}"] = d[:, 1].astype(np.float32)
            out["time"] = d[:, 0].astype(np.float32)
    out["maxnorm_step"], out["maxnorm_value"] = maxnorms(os.path.join(REF_NOABS, "output_solver.txt"))
    np.savez_compressed(os.path.join(HERE, "noabs_ref_seis.npz"), **out)
    print("wrote noabs_ref_seis.npz", len(out["time"]), "samples,", len(out["maxnorm_step"]), "max-norm lines")


def main():
    cpml()
    no_absorbing()
    attenuation()
    waterlayer()
    acoustic()
    out = {}
    for st in ("X20", "X30", "X40", "X50"):
        for ch in "XYZ":
            d = np.loadtxt(os.path.join(REF, f"DB.{st}.BX{ch}.semd"))
            out[f"{st}_{ch}"] = d[:NKEEP, 1].astype(np.float32)
            out["time"] = d[:NKEEP, 0].astype(np.float32)
    steps, norms = [], []
    txt = open(os.path.join(REF, "output_solver.txt")).read()
    for m in re.finditer(r"Time step #\s+(\d+).*?Max norm displacement vector U in all slices \(m\) =\s+([0-9.E+-]+)", txt, re.S):
        steps.append(int(m.group(1)))
        norms.append(float(m.group(2)))
    out["maxnorm_step"] = np.array(steps, dtype=np.int32)
    out["maxnorm_value"] = np.array(norms, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "halfspace_ref_seis.npz"), **out)
    print("wrote halfspace_ref_seis.npz", {k: v.shape for k, v in out.items() if k.startswith("max")})


if __name__ == "__main__":
    sys.exit(main())
