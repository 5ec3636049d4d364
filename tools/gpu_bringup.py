#!/usr/bin/env python
"""First-contact diagnostics on a B200 box: every device entry point against the CPU oracle.
Prints max abs errors; not a test (tests/ has the asserted versions)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "mcmurchie-davidson_b200"))
sys.path.insert(0, ROOT)

from oracle import oracle as O
from mmd._b200 import engine as E, synth
from mmd.molecule import Molecule

G = os.path.join(ROOT, "tests", "golden")


def t(msg, t0):
    print("   [%.2fs] %s" % (time.time() - t0, msg), flush=True)


def main():
    tf, ms = E.fp64_peak(0)
    print("FP64 DFMA peak probe: %.2f TFLOP/s (%.3f ms)" % (tf, ms), flush=True)
    # Boys
    rng = np.random.default_rng(0)
    Ts = np.concatenate([10 ** rng.uniform(-8, 4, 4000), [0.0, 39.9, 39.99999, 40.0, 40.1, 1e5], rng.uniform(0, 45, 4000)])
    for n in (0, 1, 4, 8):
        got = E.boys(n, Ts)
        ref = np.array([[O.boys(m, T) for m in range(n + 1)] for T in Ts])
        rel = np.abs(got - ref) / np.abs(ref)
        print("Boys n<=%d  max rel err %.3e  max abs err %.3e" % (n, rel.max(), np.abs(got - ref).max()), flush=True)

    for cfg in ("h2o_sto3g", "h2o_ccpvdz"):
        geom, basis = synth.config(cfg)
        t0 = time.time()
        mol = Molecule(geom, basis)
        N = mol.nbasis
        eng = mol.engine
        t("engine %s N=%d pairs=%s primpairs=%s" % (cfg, N, eng.npairs.tolist(), eng.nprimpairs.tolist()), t0)
        gold = np.load(os.path.join(G, cfg + ".npz"))
        T_ref = np.zeros((N,) * 4)
        O.doERIs(N, T_ref, mol.bfs)
        t("oracle doERIs", t0)
        T_gpu = eng.dense()
        t("gpu dense", t0)
        print(" %s TwoE max abs err vs oracle: %.3e   sum=%.12f fro=%.12f" % (cfg, np.abs(T_gpu - T_ref).max(), T_gpu.sum(), np.linalg.norm(T_gpu)), flush=True)
        # generic kernel cross-check on random quartets
        idx = rng.integers(0, N, size=(3000, 4))
        va = eng.eri_quartets(idx, impl=0)
        vb = eng.eri_quartets(idx, impl=1)
        vr = T_ref[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]]
        print(" quartets: class kernels err %.3e   generic kernel err %.3e" % (np.abs(va - vr).max(), np.abs(vb - vr).max()), flush=True)
        scr = eng.schwarz()
        print(" schwarz max abs err vs golden(ref): %.3e" % np.abs(scr.flat - gold["screen"]).max(), flush=True)
        S_, T_, V_, M_, L_ = eng.onee([a.charge for a in mol.atoms], [a.origin for a in mol.atoms], mol.center_of_charge)
        print(" onee err S %.2e T %.2e V %.2e M %.2e L %.2e" % tuple(np.abs(x - gold[k]).max() for x, k in ((S_, "S"), (T_, "T"), (V_, "V"), (M_, "M"), (L_, "L"))), flush=True)
        for tag, P, Po, Gk in (("first", gold["P1"], np.zeros_like(gold["P1"]), "G1"), ("incr", gold["Pc"], gold["Pold"], "G2"), ("cplx", gold["Pz"], np.zeros_like(gold["Pz"]), "G3")):
            Gg = eng.formPT(P, Po, screen=scr, tol=1e-12)
            Go, cnt = O.formPT(P, Po, mol.bfs, N, scr.flat, 1e-12, return_count=True)
            print(" formPT[%s] err vs golden(ref) %.3e  vs oracle %.3e   stats q=%d pq=%d fn=%d (oracle fn-quartets computed %d)" % (
                tag, np.abs(Gg - gold[Gk]).max(), np.abs(Gg - Go).max(), eng.last_stats["quartets"], eng.last_stats["prim_quartets"], eng.last_stats["fn_quartets"], cnt), flush=True)
        J, K = eng.jk_incore(gold["Pz"])
        print(" jk_incore err J %.3e K %.3e" % (np.abs(J - gold["J3"]).max(), np.abs(K - gold["K3"]).max()), flush=True)

    anchors = json.load(open(os.path.join(G, "anchors.json")))
    from mmd.postscf import PostSCF
    for name in ("h2o_sto3g_incore", "h2o_sto3g_direct", "ch4_sto3g_incore", "ch4_sto3g_direct", "h2o_ccpvdz_incore", "h2o_ccpvdz_direct"):
        if name not in anchors:
            continue
        a = anchors[name]
        geom = a.get("geometry") or synth.water()
        basis = a.get("basis") or ("sto-3g" if "sto3g" in name else "cc-pvdz")
        t0 = time.time()
        mol = Molecule(geom, basis)
        mol.RHF(doPrint=False, direct=name.endswith("direct"), conver=a.get("conver", 1e-8))
        line = " %s: E=%.12f (ref %.12f, diff %.2e) iters %s (ref %s) %.2fs" % (name, mol.energy.real, a["energy"], mol.energy.real - a["energy"], getattr(mol, "scf_iterations", None), a["iterations"], time.time() - t0)
        if "emp2" in a:
            PostSCF(mol).MP2()
            line += " MP2 diff %.2e" % (mol.emp2.real - a["emp2"])
        print(line, flush=True)


if __name__ == "__main__":
    main()
