#!/usr/bin/env python
"""Small end-to-end workload for compute-sanitizer: every kernel family once (dense fill, Schwarz, in-core J/K,
screening + direct J/K digestion incl. the per-function list and the deterministic mode, one-electron, AO->MO/MP2).
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "mcmurchie-davidson_b200"))
from mmd._b200 import synth          # noqa: E402
from mmd.molecule import Molecule    # noqa: E402
from mmd.postscf import PostSCF      # noqa: E402

geom, basis = synth.config("h2o_ccpvdz")
m = Molecule(geom, basis)
m.RHF(doPrint=False, direct=False)
PostSCF(m).MP2()
print("in-core", m.energy.real, m.emp2.real)
m = Molecule(geom, basis)
m.RHF(doPrint=False, direct=True)
print("direct ", m.energy.real)
os.environ["MMDB_DETERMINISTIC"] = "1"
m = Molecule(geom, basis)
m.RHF(doPrint=False, direct=True)
print("direct deterministic", m.energy.real)

# round 2: forces (gradient kernels), the plain-class build (flags bit 2), the warp-autonomous screen and the far lists
del os.environ["MMDB_DETERMINISTIC"]
import numpy as np  # noqa: E402
m = Molecule(geom, basis)
m.RHF(doPrint=False, direct=True)
m.forces()
print("forces", float(np.abs(np.array([a.forces for a in m.atoms])).max()))
eng = m.engine
scr = eng.schwarz()
P = np.asarray(m.P, dtype=complex)
Z = np.zeros_like(P)
G0 = eng.formPT(P, Z, screen=scr, tol=1e-12)
G1 = eng.formPT(P, Z, screen=scr, tol=1e-12, flags=4)
os.environ["MMDB_SCREEN_WARPS"] = "1"
G2 = eng.formPT(P, Z, screen=scr, tol=1e-12)
del os.environ["MMDB_SCREEN_WARPS"]
os.environ["MMDB_FAR_MAXL"] = "3"
G3 = eng.formPT(P, Z, screen=scr, tol=1e-12, flags=4)
print("grouped vs plain / warp screen / far lists", float(np.abs(G0 - G1).max()), float(np.abs(G0 - G2).max()), float(np.abs(G0 - G3).max()))
