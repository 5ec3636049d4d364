"""Direct (integral-driven) Fock build — same entry point as the reference's cython/fock.pyx.

    formPT(P, P_old, bfs, nbasis, screen, tol) -> (N,N) complex128, UN-symmetrised
"""
import numpy as np

from mmd._b200 import engine as _engine


def formPT(P, P_old, bfs, nbasis, screen, tol):
    P = np.asarray(P)
    P_old = np.asarray(P_old)
    # the reference's typed signature (np.ndarray[complex, ndim=2]) rejects anything but complex128
    for name, arr in (("P", P), ("P_old", P_old)):
        if arr.dtype != np.complex128:
            raise ValueError("Buffer dtype mismatch, expected 'complex' but got '%s' for %s" % (arr.dtype, name))
        if arr.ndim != 2:
            raise ValueError("Buffer has wrong number of dimensions (expected 2, got %d)" % arr.ndim)
    N = int(nbasis)
    bfs = list(bfs)
    if len(bfs) != N or P.shape != (N, N) or P_old.shape != (N, N):
        raise ValueError("formPT: inconsistent nbasis / matrix shapes")
    eng = _engine.engine_for(bfs)
    # no statistics on the reference-facing call: they cost four more atomics per screening tile, a read-back and a
    # stream synchronisation in the middle of the build (Engine.formPT(..., want_stats=True) keeps them for the tools)
    return eng.formPT(P, P_old, screen=screen, tol=float(tol), want_stats=False)
