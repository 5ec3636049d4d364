#!/usr/bin/env python
"""Driver for ncu: dense fill + in-core J/K of benzene/6-31G** (config 2)."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "mcmurchie-davidson_b200"))
from mmd._b200 import synth          # noqa: E402
from mmd.molecule import Molecule    # noqa: E402

mol = Molecule(*synth.config("benzene_631gss"))
eng = mol.engine
T = eng.dense()
rng = np.random.default_rng(0)
A = rng.standard_normal((mol.nbasis, mol.nbasis))
for _ in range(3):
    J, K = eng.jk_incore(A + A.T)
print(float(J.real.sum()), float(K.real.sum()))
