#!/usr/bin/env python
"""Small driver for ncu: W warm-up direct Fock builds + 1 profiled build of a workload.
    ncu ... python tools/profile_direct.py w8_ccpvdz"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "mcmurchie-davidson_b200"))
from mmd._b200 import synth          # noqa: E402
from mmd.molecule import Molecule    # noqa: E402


def main(workload, builds):
    import scipy.linalg
    mol = Molecule(*synth.config(workload))
    mol.one_electron_integrals()
    FO = mol.X.T @ mol.Core @ mol.X
    _, CO = scipy.linalg.eigh(FO)
    Cm = mol.X @ CO
    P = (Cm[:, :mol.nocc] @ Cm[:, :mol.nocc].conj().T).astype(complex)
    eng = mol.engine
    scr = eng.schwarz()
    for _ in range(builds):
        eng.formPT(P, np.zeros_like(P), screen=scr, tol=1e-12)
    print(eng.last_stats["quartets"], eng.last_stats["prim_quartets"])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "w8_ccpvdz", int(sys.argv[2]) if len(sys.argv) > 2 else 2)
