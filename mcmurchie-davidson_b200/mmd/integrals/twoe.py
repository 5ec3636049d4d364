"""Two-electron integrals — same names and call signatures as the reference's Cython module
(cython/twoe.pyx + cython/basis.pxi), evaluated by the CUDA shell-quartet kernels.

    Basis(origin, shell, num_exps, exps, coefs)      cython/basis.pxi:46
    ERI(a, b, c, d) -> float                         cython/twoe.pyx:36-50
    doERIs(N, TwoE, bfs) -> (N,N,N,N) float64        cython/twoe.pyx:12-31
"""
import math

import numpy as np

from mmd._b200 import engine as _engine


def _dfact(n):
    """n!! with (-1)!! = 0!! = 1 — the semantics the reference's normalisation relies on."""
    out = 1.0
    while n > 1:
        out *= n
        n -= 2
    return out


class Basis(object):
    """One contracted Cartesian Gaussian basis function (host-side carrier).

    Read-only attributes as in the reference: origin (3,), shell (3,) int64 = (l,m,n), num_exps,
    exps (K,), coefs (K,) — the contraction coefficients AFTER contracted normalisation — and
    norm (K,), the primitive normalisation constants (cython/basis.pxi:87-120).
    """

    __slots__ = ("_origin", "_shell", "_num_exps", "_exps", "_coefs", "_norm", "_raw_coefs", "__weakref__")

    def __init__(self, origin, shell, num_exps, exps, coefs):
        self._origin = np.array([float(origin[k]) for k in range(3)], dtype=np.float64)
        self._shell = np.array([int(shell[k]) for k in range(3)], dtype=np.int64)
        self._num_exps = int(num_exps)
        self._exps = np.array([float(exps[k]) for k in range(self._num_exps)], dtype=np.float64)
        self._raw_coefs = np.array([float(coefs[k]) for k in range(self._num_exps)], dtype=np.float64)
        self._normalize()

    def _normalize(self):
        l, m, n = (int(x) for x in self._shell)
        lam = l + m + n
        ff = _dfact(2 * l - 1) * _dfact(2 * m - 1) * _dfact(2 * n - 1)
        # primitive norms
        self._norm = np.sqrt(np.power(2.0, 2 * lam + 1.5) * np.power(self._exps, lam + 1.5) / ff / math.pi ** 1.5)
        # contracted norm: <phi|phi> = pi^1.5 ff / 2^lam * sum_ab N_a N_b d_a d_b / (a+b)^(lam+1.5)
        w = self._norm * self._raw_coefs
        pair = np.add.outer(self._exps, self._exps) ** (lam + 1.5)
        total = float(np.sum(np.outer(w, w) / pair)) * (math.pi ** 1.5) * ff / (2.0 ** lam)
        self._coefs = self._raw_coefs * total ** -0.5

    origin = property(lambda self: self._origin.copy())
    shell = property(lambda self: self._shell.copy())
    num_exps = property(lambda self: self._num_exps)
    exps = property(lambda self: self._exps.copy())
    coefs = property(lambda self: self._coefs.copy())
    norm = property(lambda self: self._norm.copy())

    def __repr__(self):
        return "Basis(origin=%s, shell=%s, K=%d)" % (self._origin.tolist(), self._shell.tolist(), self._num_exps)


def _require_basis(*objs):
    for o in objs:
        if not isinstance(o, Basis):
            raise TypeError("Argument has incorrect type (expected mmd.integrals.twoe.Basis, got %s)" % type(o).__name__)


def ERI(a, b, c, d):
    """Contracted electron repulsion integral (ab|cd), chemists' notation."""
    _require_basis(a, b, c, d)
    uniq = []
    pos = []
    for x in (a, b, c, d):
        for k, y in enumerate(uniq):
            if y is x:
                pos.append(k)
                break
        else:
            uniq.append(x)
            pos.append(len(uniq) - 1)
    eng, where = _engine.engine_containing(uniq)       # a molecule's (or a cached) engine that already holds them
    if eng is not None:
        return float(eng.eri_quartets(np.array([[where[k] for k in pos]]))[0])
    eng = _engine.engine_for(uniq)
    return float(eng.eri_quartets(np.array([pos]))[0])


def doERIs(N, TwoE, bfs):
    """Fill the dense (N,N,N,N) tensor with all eight permutational images of every unique integral.
    Writes into the caller's buffer (like the reference) and returns it."""
    N = int(N)
    bfs = list(bfs)
    if len(bfs) != N:
        raise ValueError("doERIs: N does not match len(bfs)")
    _require_basis(*bfs)
    eng = _engine.engine_for(bfs)
    host = eng.dense(keep_device=True)
    out = np.asarray(TwoE)
    if out.shape != (N, N, N, N):
        raise ValueError("doERIs: TwoE must have shape (N,N,N,N)")
    out[...] = host
    return TwoE
