#!/usr/bin/env python
"""Large stratified ERI samples of the benchmark configurations, values FROM THE REFERENCE ITSELF
(oracle/_ref = the reference's compiled Cython `ERI`, cython/twoe.pyx:36-50), SURVEY.md 8(d): ">= 1e4 sampled
quartets per class" for configurations 2-5.  Build-container only:

    python tests/golden/make_golden_sampled.py [config ...]      # writes tests/golden/sampled_big_<config>.npz

Strata: the 21 canonical angular-momentum classes (la lb|lc ld) x five bins of the distance between the two pair
centres (same centre, (0,3], (3,8], (8,16], > 16 bohr), so every class has primitive quartets in every Boys regime
(tabulated range, the 37 <= T < 60 switch-over window, pure asymptotic).  Pairs are drawn among the significant
ones (sqrt(pq|pq) > 1e-5).  Stored: function quartets as uint16, values as float64, class / bin ids as uint8.
"""
import importlib.util
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from mmd.molecule import Molecule            # noqa: E402  (the REFERENCE)
from mmd.integrals.twoe import ERI           # noqa: E402

assert os.path.realpath(sys.modules["mmd.molecule"].__file__).startswith(os.path.realpath(REF))
spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "mcmurchie-davidson_b200", "mmd", "_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)
from oracle import oracle as O               # noqa: E402  (C restatement: Schwarz table to choose samples only)

BINS = [0.0, 3.0, 8.0, 16.0]
_BFS = None


def _work(chunk):
    return [ERI(_BFS[i], _BFS[j], _BFS[k], _BFS[l]) for i, j, k, l in chunk]


def sample(name, nper):
    global _BFS
    mol = Molecule(geometry=synth.config(name)[0], basis=synth.config(name)[1])
    _BFS = mol.bfs
    N = mol.nbasis
    fb = O.FlatBasis(mol.bfs)
    sq = np.sqrt(np.abs(O.schwarz(fb)))
    L = np.asarray([int(sum(b.shell)) for b in mol.bfs])
    C = np.asarray([np.asarray(b.origin, dtype=np.float64) for b in mol.bfs])
    p, q = np.tril_indices(N)
    mid = 0.5 * (C[p] + C[q])
    rng = np.random.default_rng(11)
    pcs = [(la, lb) for la in range(3) for lb in range(la + 1)]
    sel = {}
    for la, lb in pcs:
        m = ((L[p] == la) & (L[q] == lb)) | ((L[p] == lb) & (L[q] == la))
        sel[(la, lb)] = np.nonzero(m & (sq > 1e-5))[0]
    idx, cls, bins = [], [], []
    nclass = 0
    for ib, (la, lb) in enumerate(pcs):
        for (lc, ld) in pcs[:ib + 1]:
            pa, pb = sel[(la, lb)], sel[(lc, ld)]
            if len(pa) == 0 or len(pb) == 0:
                continue
            got = 0
            per_bin = nper // 5
            for b in range(5):
                # rejection sampling into the distance bin
                want = per_bin if b < 4 else nper - got
                have = 0
                for _ in range(60):
                    x = rng.choice(pa, size=4 * want)
                    y = rng.choice(pb, size=4 * want)
                    d = np.linalg.norm(mid[x] - mid[y], axis=1)
                    if b == 0:
                        ok = d == 0.0
                    elif b < 4:
                        ok = (d > BINS[b - 1]) & (d <= BINS[b])
                    else:
                        ok = d > BINS[3]
                    x, y = x[ok][:want - have], y[ok][:want - have]
                    for xa, ya in zip(x, y):
                        # random bra/ket order and random order inside the pairs: exercises every orientation
                        a, bq, c, dq = p[xa], q[xa], p[ya], q[ya]
                        if rng.random() < 0.5:
                            a, bq = bq, a
                        if rng.random() < 0.5:
                            c, dq = dq, c
                        if rng.random() < 0.5:
                            a, bq, c, dq = c, dq, a, bq
                        idx.append((a, bq, c, dq)); cls.append(nclass); bins.append(b)
                    have += len(x)
                    if have >= want:
                        break
                got += have
            # top up from the unconstrained distribution if some bin was empty (small molecules)
            while got < nper:
                xa, ya = rng.choice(pa), rng.choice(pb)
                idx.append((p[xa], q[xa], p[ya], q[ya])); cls.append(nclass); bins.append(5)
                got += 1
            nclass += 1
    idx = np.asarray(idx, dtype=np.int64)
    t0 = time.time()
    nproc = os.cpu_count() or 1
    chunks = [idx[i:i + 2000].tolist() for i in range(0, len(idx), 2000)]
    with mp.get_context("fork").Pool(nproc) as pool:
        res = pool.map(_work, chunks)
    vals = np.asarray([v for r in res for v in r])
    dt = time.time() - t0
    np.savez_compressed(os.path.join(HERE, "sampled_big_%s.npz" % name), idx=idx.astype(np.uint16), vals=vals,
                        cls=np.asarray(cls, dtype=np.uint8), bins=np.asarray(bins, dtype=np.uint8), N=N,
                        ref_wall_seconds=dt, ref_processes=nproc)
    print("sampled_big", name, len(idx), "function quartets in", nclass, "classes; reference ERI wall %.1f s on %d processes" % (dt, nproc), flush=True)


if __name__ == "__main__":
    cfgs = sys.argv[1:] or ["benzene_631gss", "w8_ccpvdz", "c20h42_631gs", "w32_ccpvdz"]
    for c in cfgs:
        sample(c, 10000 if c in ("c20h42_631gs", "w32_ccpvdz") else 2500)
