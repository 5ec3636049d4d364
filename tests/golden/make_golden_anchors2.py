#!/usr/bin/env python
"""More anchors FROM THE REFERENCE ITSELF (oracle/_ref), for the smoke configurations of the reference's tests that
round 1 did not pin on the GPU path, and BASELINE config 2.  Build container only:

    python tests/golden/make_golden_anchors2.py [--benzene]      # writes tests/golden/anchors2.json, updatefock_h2o.npz

  he_ccpvtz_incore        reference tests/test007.py
  h2co_sto3g_incore       reference tests/test008.py (energy and dipole)
  updatefock (npz)        a genuinely complex Hermitian density pushed through mol.updateFock() — the call real-time
                          propagation makes every step (mmd/realtime.py:62) and the reason the J/K boundary is complex
  benzene_631gss_incore   BASELINE config 2 (--benzene: ~40 minutes of single-threaded Cython doERIs in the reference)
"""
import importlib.util
import io
import json
import os
import sys
import time
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, REF)
from mmd.molecule import Molecule            # noqa: E402  (the REFERENCE)

assert os.path.realpath(sys.modules["mmd.molecule"].__file__).startswith(os.path.realpath(REF))
spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "mcmurchie-davidson_b200", "mmd", "_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

HELIUM = "\n0 1\nHe    0.000000    0.000000    0.000000\n"
H2CO = """
0 1
C          0.0000000000        0.0000000000       -0.5265526741
O          0.0000000000        0.0000000000        0.6555124750
H          0.0000000000       -0.9325664988       -1.1133424527
H          0.0000000000        0.9325664988       -1.1133424527
"""


def run(geom, basis, direct=False):
    mol = Molecule(geometry=geom, basis=basis)
    buf = io.StringIO()
    t0 = time.time()
    with redirect_stdout(buf):
        mol.RHF(direct=direct)
    txt = buf.getvalue()
    its = int(txt.split(" in ")[1].split()[0])
    return mol, {"geometry": geom, "basis": basis, "direct": direct, "energy": float(mol.energy.real), "iterations": its,
                 "dipole": [float(x.real) for x in mol.mu], "P_RMS": float(np.real(mol.P_RMS)), "ref_seconds": time.time() - t0}


def main():
    path = os.path.join(HERE, "anchors2.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    if "--benzene" in sys.argv:
        geom, basis = synth.config("benzene_631gss")
        _, out["benzene_631gss_incore"] = run(geom, basis)
        print("benzene", out["benzene_631gss_incore"]["energy"], out["benzene_631gss_incore"]["iterations"], flush=True)
    else:
        _, out["he_ccpvtz_incore"] = run(HELIUM, "cc-pvtz")
        _, out["h2co_sto3g_incore"] = run(H2CO, "sto-3g")
        mol, _ = run(synth.water(), "sto-3g")
        # complex Hermitian density in the orthonormal basis: PO -> unOrthoDen -> buildFock (complex einsum J/K) -> orthoFock
        rng = np.random.default_rng(5)
        N = mol.nbasis
        mol.orthoDen()
        A = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
        PO = mol.PO + 0.05 * (A + A.conj().T)
        mol.PO = PO.copy()
        mol.updateFock()
        np.savez_compressed(os.path.join(HERE, "updatefock_h2o.npz"), PO=PO, P=mol.P, F=mol.F, FO=mol.FO, J=mol.J, K=mol.K)
        print({k: (v["energy"], v["iterations"]) for k, v in out.items()})
    json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
