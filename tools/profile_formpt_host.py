#!/usr/bin/env python
"""Where the host-buffer formPT call spends its time beyond the device build (e2e vs value in bench.py)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "mcmurchie-davidson_b200"))
from mmd._b200 import synth          # noqa: E402
from mmd.integrals.fock import formPT   # noqa: E402
from mmd.molecule import Molecule    # noqa: E402


def main(workload):
    import cProfile
    import pstats
    import scipy.linalg
    import torch
    mol = Molecule(*synth.config(workload))
    mol.one_electron_integrals()
    FO = mol.X.T @ mol.Core @ mol.X
    _, CO = scipy.linalg.eigh(FO)
    Cm = mol.X @ CO
    P = (Cm[:, :mol.nocc] @ Cm[:, :mol.nocc].conj().T).astype(complex)
    Z = np.zeros_like(P)
    eng = mol.engine
    scr = eng.schwarz()
    for _ in range(3):
        formPT(P, Z, mol.bfs, mol.nbasis, scr, 1e-12)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        formPT(P, Z, mol.bfs, mol.nbasis, scr, 1e-12)
    print("formPT host-buffer call: %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        formPT(P, Z, mol.bfs, mol.nbasis, scr, 1e-12)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "w32_ccpvdz")
