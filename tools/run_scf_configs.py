#!/usr/bin/env python
"""End-to-end RHF on the benchmark configurations through the drop-in API (Molecule(...).RHF()).
Prints energy, iterations and wall time per configuration; direct and in-core energies must agree."""
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "mcmurchie-davidson_b200"))
from mmd._b200 import synth          # noqa: E402
from mmd.molecule import Molecule    # noqa: E402

runs = [("benzene_631gss", False), ("benzene_631gss", True), ("w8_ccpvdz", False), ("w8_ccpvdz", True),
        ("c20h42_631gs", True), ("w32_ccpvdz", True)]
if len(sys.argv) > 1:
    runs = [r for r in runs if r[0] in sys.argv[1:]]
for cfg, direct in runs:
    t0 = time.time()
    mol = Molecule(*synth.config(cfg))
    mol.RHF(doPrint=False, direct=direct)
    dt = time.time() - t0
    print("%-16s %-8s N=%4d  E(RHF) = %.10f  iterations %3s  converged %s  wall %.1f s" % (
        cfg, "direct" if direct else "in-core", mol.nbasis, mol.energy.real, getattr(mol, "scf_iterations", None),
        mol.is_converged, dt), flush=True)
    for it, (q, c, ms) in enumerate(getattr(mol, "fock_trace", []) or []):
        print("    build %2d: %11d quartets of %11d candidates  %8.3f ms" % (it + 1, q, c, ms), flush=True)
