#!/usr/bin/env python
"""Golden Fock elements at the benchmark sizes (configurations 4 and 5), generated with the PINNED
C oracle (oracle/md_oracle.c, threads) — run in the build container:

    python tests/golden/make_golden_fock.py            # writes tests/golden/fock_elements_<config>.npz

For each configuration and for two closed-form densities (bit-reproducible from the geometry alone, so the
fixture need not carry an N x N matrix) a stratified set of elements (p, q) of

    sym(G)_pq = 1/2 (G + G^T)_pq = sum_rs P_sr [ 2 (pq|rs) - (ps|qr) ]          (mmd/scf.py:93,97-99)

is summed integral by integral with the oracle's ERI (cython/twoe.pyx:36-50 restated).  Integrals whose Schwarz
bound times |P_sr| is below 1e-19 are skipped (at most N^2 = 6.4e5 terms per element: error < 1e-13).
The GPU test compares formPT at tol = 0 and tol = 1e-12 with these values (<= 1e-10).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "mcmurchie-davidson_b200"))
sys.path.insert(0, ROOT)

from mmd._b200 import synth                 # noqa: E402
from mmd.molecule import Molecule           # noqa: E402
from oracle import oracle as O              # noqa: E402

CUT = 1e-19


def densities(bfs):
    """Two closed-form real symmetric 'densities' (tests/test_gpu_parity.py rebuilds them from the geometry):
    A: dense, oscillating, O(0.1) everywhere (nothing is density-screened);
    B: local, decaying with the distance between the function centres (mimics a real density; the
       density-weighted screen of cython/fock.pyx:46-57 removes most far quartets)."""
    N = len(bfs)
    C = np.array([np.asarray(b.origin, dtype=np.float64) for b in bfs])
    i = np.arange(N, dtype=np.float64)
    A = 0.1 * np.cos(0.37 * (i[:, None] + i[None, :])) + 0.05 * np.cos(0.011 * (i[:, None] - i[None, :]) ** 2)
    A = 0.5 * (A + A.T)
    r2 = ((C[:, None, :] - C[None, :, :]) ** 2).sum(-1)
    B = np.exp(-0.35 * r2) * (0.3 * np.cos(0.61 * (i[:, None] + i[None, :])) + 0.2)
    B = 0.5 * (B + B.T)
    return {"A": A, "B": B}


def pick_elements(bfs, Qfull, n=48, seed=7):
    """Stratified (p, q): diagonal / same-atom / neighbour / far pairs, s, p and d functions."""
    rng = np.random.default_rng(seed)
    N = len(bfs)
    L = np.array([int(np.sum(b.shell)) for b in bfs])
    C = np.array([np.asarray(b.origin, dtype=np.float64) for b in bfs])
    out = []
    want = [(0, 0), (1, 0), (1, 1), (2, 0), (2, 1), (2, 2)]
    for la, lb in want:
        pa = np.nonzero(L == la)[0]
        pb = np.nonzero(L == lb)[0]
        for kind in list(range(4)) * 2:
            for _ in range(400):
                p = int(rng.choice(pa))
                q = int(rng.choice(pb))
                d = np.linalg.norm(C[p] - C[q])
                ok = [(p == q) if la == lb else d == 0.0, d == 0.0 and p != q, 0.0 < d < 4.5, d >= 4.5 and Qfull[p, q] > 1e-9][kind]
                if ok:
                    out.append((max(p, q), min(p, q)))
                    break
    out = sorted(set(out))
    return np.array(out[:n] if len(out) > n else out, dtype=np.int64)


def element(fb, N, Qfull, P, p, q):
    """sum_rs P_sr [2 (pq|rs) - (ps|qr)] with Schwarz skipping below CUT."""
    absP = np.abs(P)
    # Coulomb: pairs (r,s), r >= s, weight P_sr + P_rs (r != s)
    r, s = np.tril_indices(N)
    w = np.where(r == s, P[r, s], P[r, s] + P[s, r])
    keep = np.sqrt(np.abs(Qfull[p, q] * Qfull[r, s])) * np.abs(w) >= CUT
    r, s, w = r[keep], s[keep], w[keep]
    idx = np.stack([np.full_like(r, p), np.full_like(r, q), r, s], axis=1)
    J = float(np.dot(O.ERI_batch(fb, idx), w))
    # exchange: all (s, r): (ps|qr) P_sr
    s2, r2 = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    s2 = s2.ravel(); r2 = r2.ravel()
    keep = np.sqrt(np.abs(Qfull[p, s2] * Qfull[q, r2])) * absP[s2, r2] >= CUT
    s2, r2 = s2[keep], r2[keep]
    idx = np.stack([np.full_like(s2, p), s2, np.full_like(s2, q), r2], axis=1)
    K = float(np.dot(O.ERI_batch(fb, idx), P[s2, r2]))
    return 2.0 * J - K, len(r) + len(s2)


def main(configs):
    for cfg in configs:
        t0 = time.time()
        mol = Molecule(*synth.config(cfg))
        N = mol.nbasis
        fb = O.FlatBasis(mol.bfs)
        flat = O.schwarz(fb)
        Qfull = np.zeros((N, N))
        a, b = np.tril_indices(N)
        Qfull[a, b] = flat
        Qfull[b, a] = flat
        el = pick_elements(mol.bfs, Qfull)
        dens = densities(mol.bfs)
        out = {"elements": el, "nbasis": N}
        for name, P in dens.items():
            vals = np.zeros(len(el))
            nint = 0
            for n, (p, q) in enumerate(el):
                vals[n], k = element(fb, N, Qfull, P, int(p), int(q))
                nint += k
                print("%s %s (%d,%d) = %.15e   [%d integrals, %.0f s]" % (cfg, name, p, q, vals[n], k, time.time() - t0), flush=True)
            out["G_" + name] = vals
            out["nint_" + name] = nint
        np.savez_compressed(os.path.join(HERE, "fock_elements_%s.npz" % cfg), **out)
        print(cfg, "done in %.0f s" % (time.time() - t0))


if __name__ == "__main__":
    main(sys.argv[1:] or ["c20h42_631gs", "w32_ccpvdz"])
