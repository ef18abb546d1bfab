"""Regenerates tests/golden/*.npz from the reference tree (run in the build container, where /root/reference exists).

halfspace_ref_seis.npz : the reference's own golden seismograms for the default example (config 1),
    EXAMPLES/applications/homogeneous_halfspace/REF_SEIS/DB.X{20,30,40,50}.BX{X,Y,Z}.semd, first NKEEP samples
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
NKEEP = 1500
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


def main():
    cpml()
    attenuation()
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
