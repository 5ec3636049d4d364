#!/usr/bin/env python
"""Draw the bounded, stratified sample of surviving shell quartets that bench.py's cpu_baseline /
--impl reference legs time with the reference's own Cython ERI.

Run on a GPU box (uses the engine for the Schwarz table and the class populations of a real direct
build); writes bench_samples/<workload>.json (committed), e.g.

    gpurun -- 'python tools/make_bench_sample.py w32_ccpvdz 40 && cp bench_samples/*.json gpurun_out/'

Strata = angular-momentum classes (bra pair class | ket pair class).  Within a class, shell quartets
are drawn uniformly from the survivors of the shell-level bound (rejection sampling), so the
contraction-depth mix inside the class is the workload's own.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "mcmurchie-davidson_b200"))
from mmd._b200 import lib as L, synth          # noqa: E402
from mmd.molecule import Molecule              # noqa: E402

NAMES = ["ss", "ps", "pp", "ds", "dp", "dd"]


def main(workload, per_class):
    import scipy.linalg
    geom, basis = synth.config(workload)
    mol = Molecule(geom, basis)
    N = mol.nbasis
    eng = mol.engine
    t = eng.table
    mol.one_electron_integrals()
    FO = mol.X.T @ mol.Core @ mol.X
    _, CO = scipy.linalg.eigh(FO)
    Cm = mol.X @ CO
    P = (Cm[:, :mol.nocc] @ Cm[:, :mol.nocc].conj().T).astype(complex)
    scr = eng.schwarz()
    tol = 1e-12
    eng.formPT(P, np.zeros_like(P), screen=scr, tol=tol)
    pops = {k: int(v["quartets"]) for k, v in eng.last_stats["classes"].items()}
    # shell-pair bounds and shell-block density maxima on the host
    Q = np.zeros((N, N))
    p, q = np.tril_indices(N)
    Q[p, q] = scr.flat
    Q[q, p] = scr.flat
    SQ = np.sqrt(np.abs(Q))
    ns = t.nshell
    nf = np.array([(l + 1) * (l + 2) // 2 for l in t.am])
    Qs = np.zeros((ns, ns))
    DS = np.zeros((ns, ns))
    D = np.abs(P)
    for A in range(ns):
        for B in range(ns):
            sa = slice(t.bf0[A], t.bf0[A] + nf[A])
            sb = slice(t.bf0[B], t.bf0[B] + nf[B])
            Qs[A, B] = SQ[sa, sb].max()
            DS[A, B] = D[sa, sb].max()
    pairs = {}
    for pc in range(L.NCLASS_PAIR):
        A, B = np.nonzero((eng.pair_class == pc) & ~eng.pair_flip)
        keep = eng.pair_index[A, B] >= 0
        # unique unordered pairs: stored orientation only
        sel = [(a, b) for a, b in zip(A[keep], B[keep]) if not (a != b and eng.pair_flip[a, b])]
        uniq = {}
        for a, b in sel:
            uniq[int(eng.pair_index[a, b])] = (int(a), int(b))
        pairs[pc] = [uniq[k] for k in sorted(uniq)]
    rng = np.random.default_rng(0)
    out = []
    for cb in range(L.NCLASS_PAIR):
        for ck in range(cb + 1):
            key = "(%s|%s)" % (NAMES[cb], NAMES[ck])
            if pops.get(key, 0) == 0:
                continue
            got, tries = 0, 0
            while got < per_class and tries < 200000:
                tries += 1
                ib = int(rng.integers(len(pairs[cb])))
                ik = int(rng.integers(len(pairs[ck])))
                if cb == ck and ik > ib:
                    continue
                A, B = pairs[cb][ib]
                Cc, Dd = pairs[ck][ik]
                dmax = max(4 * DS[A, B], 4 * DS[Cc, Dd], DS[A, Cc], DS[A, Dd], DS[B, Cc], DS[B, Dd])
                if Qs[A, B] * Qs[Cc, Dd] * dmax < tol:
                    continue
                fns = set()
                for a in range(nf[A]):
                    for b in range(nf[B]):
                        for c in range(nf[Cc]):
                            for d in range(nf[Dd]):
                                i, j, k, l = t.bf0[A] + a, t.bf0[B] + b, t.bf0[Cc] + c, t.bf0[Dd] + d
                                if i < j: i, j = j, i
                                if k < l: k, l = l, k
                                if i * (i + 1) // 2 + j < k * (k + 1) // 2 + l: i, j, k, l = k, l, i, j
                                fns.add((int(i), int(j), int(k), int(l)))
                out.append([key, sorted(fns)])
                got += 1
    spec = {"workload": workload, "workload_desc": "%s direct RHF Fock build, N=%d, first-iteration density, tol 1e-12" % (workload, N),
            "geometry": geom, "basis": basis, "N": N, "class_quartets": pops, "per_class": per_class, "quartets": out}
    os.makedirs(os.path.join(ROOT, "bench_samples"), exist_ok=True)
    path = os.path.join(ROOT, "bench_samples", workload + ".json")
    with open(path, "w") as f:
        json.dump(spec, f, separators=(",", ":"))
    print("wrote", path, "shell quartets:", len(out), "integrals:", sum(len(x[1]) for x in out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "w32_ccpvdz", int(sys.argv[2]) if len(sys.argv) > 2 else 40)
