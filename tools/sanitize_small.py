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
