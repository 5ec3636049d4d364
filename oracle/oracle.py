"""ctypes front-end of oracle/libmdoracle.so (md_oracle.c) — TEST INFRASTRUCTURE ONLY.

Parity status: pinned (see md_oracle.c header).  Functions mirror the reference's names:
boys/E/R (cython/util.pxi), ERI/doERIs (cython/twoe.pyx), formPT (cython/fock.pyx),
jk_incore (mmd/scf.py:97-98), normalize (cython/basis.pxi:87-120).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "libmdoracle.so")
    src = os.path.join(_HERE, "md_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libmdoracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        basis = [_dp, _lp, _lp, _lp, _dp, _dp, _dp]
        L.mdo_boys.restype = C.c_double
        L.mdo_boys.argtypes = [C.c_double, C.c_double]
        L.mdo_E.restype = C.c_double
        L.mdo_E.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.mdo_R.restype = C.c_double
        L.mdo_R.argtypes = [C.c_int] * 4 + [C.c_double] * 5
        L.mdo_electron_repulsion.restype = C.c_double
        L.mdo_electron_repulsion.argtypes = [C.c_double, _lp, _dp] * 4
        L.mdo_normalize.restype = None
        L.mdo_normalize.argtypes = [_lp, C.c_long, _dp, _dp, _dp]
        L.mdo_ERI.restype = C.c_double
        L.mdo_ERI.argtypes = [C.c_long] + basis + [C.c_long] * 4
        L.mdo_ERI_batch.restype = None
        L.mdo_ERI_batch.argtypes = [C.c_long] + basis + [C.c_long, _lp, _dp]
        L.mdo_doERIs.restype = None
        L.mdo_doERIs.argtypes = [C.c_long, _dp] + basis
        L.mdo_schwarz.restype = None
        L.mdo_schwarz.argtypes = [C.c_long, _dp] + basis
        L.mdo_formPT.restype = None
        L.mdo_formPT.argtypes = [C.c_long, _dp, _dp, _dp, C.c_double, _dp, C.POINTER(C.c_long)] + basis
        L.mdo_onee.restype = None
        L.mdo_onee.argtypes = [C.c_long, C.c_long, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp] + basis
        L.mdo_jk_incore.restype = None
        L.mdo_jk_incore.argtypes = [C.c_long, _dp, _dp, _dp, _dp]
        L.mdo_ERIx_batch.restype = None
        L.mdo_ERIx_batch.argtypes = [C.c_long] + basis + [C.c_long, _lp, _lp, _dp]
        L.mdo_onee_x.restype = C.c_double
        L.mdo_onee_x.argtypes = [C.c_long] + basis + [C.c_long] * 5
        L.mdo_V_x.restype = C.c_double
        L.mdo_V_x.argtypes = [C.c_long] + basis + [C.c_long, C.c_long, _dp, C.c_long, C.c_long]
        _LIB = L
    return _LIB


def boys(m, T):
    return lib().mdo_boys(float(m), float(T))


def E(i, j, t, Qx, a, b):
    return lib().mdo_E(i, j, t, Qx, a, b)


def R(t, u, v, n, p, PCx, PCy, PCz, RPC):
    return lib().mdo_R(t, u, v, n, p, PCx, PCy, PCz, RPC)


def normalize(shell, exps, coefs):
    """-> (normalised coefs, primitive norms), cython/basis.pxi:87-120."""
    lmn = np.ascontiguousarray(shell, dtype=np.int64)
    e = np.ascontiguousarray(exps, dtype=np.float64)
    c = np.array(coefs, dtype=np.float64)
    n = np.zeros_like(e)
    lib().mdo_normalize(lmn, len(e), e, c, n)
    return c, n


class FlatBasis(object):
    """Flat-array image of a list of reference-style Basis objects (duck-typed:
    .origin .shell .num_exps .exps .coefs(normalised) .norm)."""

    def __init__(self, bfs):
        self.nbf = len(bfs)
        self.origin = np.ascontiguousarray([np.asarray(b.origin, dtype=np.float64) for b in bfs]).reshape(-1)
        self.shell = np.ascontiguousarray([np.asarray(b.shell, dtype=np.int64) for b in bfs]).reshape(-1)
        self.nprim = np.ascontiguousarray([int(b.num_exps) for b in bfs], dtype=np.int64)
        self.off = np.zeros(self.nbf, dtype=np.int64)
        self.off[1:] = np.cumsum(self.nprim)[:-1]
        self.exps = np.ascontiguousarray(np.concatenate([np.asarray(b.exps, dtype=np.float64) for b in bfs]))
        self.coefs = np.ascontiguousarray(np.concatenate([np.asarray(b.coefs, dtype=np.float64) for b in bfs]))
        self.norm = np.ascontiguousarray(np.concatenate([np.asarray(b.norm, dtype=np.float64) for b in bfs]))

    def args(self):
        return (self.origin, self.shell, self.nprim, self.off, self.exps, self.coefs, self.norm)


def _fb(bfs):
    return bfs if isinstance(bfs, FlatBasis) else FlatBasis(bfs)


def ERI(a, b, c, d):
    fb = FlatBasis([a, b, c, d])
    return lib().mdo_ERI(4, *fb.args(), 0, 1, 2, 3)


def ERI_batch(bfs, idx):
    fb = _fb(bfs)
    idx = np.ascontiguousarray(idx, dtype=np.int64).reshape(-1, 4)
    out = np.zeros(len(idx))
    lib().mdo_ERI_batch(fb.nbf, *fb.args(), len(idx), idx.reshape(-1), out)
    return out


def doERIs(N, TwoE, bfs):
    fb = _fb(bfs)
    assert TwoE.shape == (N, N, N, N) and TwoE.flags.c_contiguous and TwoE.dtype == np.float64
    lib().mdo_doERIs(N, TwoE.reshape(-1), *fb.args())
    return TwoE


def schwarz(bfs):
    fb = _fb(bfs)
    N = fb.nbf
    out = np.zeros(N * (N + 1) // 2)
    lib().mdo_schwarz(N, out, *fb.args())
    return out


def formPT(P, P_old, bfs, nbasis, screen, tol, return_count=False):
    """screen: dict {pq: value} like the reference, or a flat array of N(N+1)/2."""
    fb = _fb(bfs)
    N = int(nbasis)
    if isinstance(screen, dict):
        scr = np.array([screen[k] for k in range(N * (N + 1) // 2)], dtype=np.float64)
    else:
        scr = np.ascontiguousarray(screen, dtype=np.float64)
    Pc = np.ascontiguousarray(P, dtype=np.complex128)
    Po = np.ascontiguousarray(P_old, dtype=np.complex128)
    G = np.zeros((N, N), dtype=np.complex128)
    cnt = C.c_long(0)
    lib().mdo_formPT(N, Pc.view(np.float64).reshape(-1), Po.view(np.float64).reshape(-1), scr, float(tol),
                     G.view(np.float64).reshape(-1), C.byref(cnt), *fb.args())
    return (G, cnt.value) if return_count else G


def jk_incore(TwoE, P):
    N = TwoE.shape[0]
    T = np.ascontiguousarray(TwoE, dtype=np.float64)
    Pc = np.ascontiguousarray(P, dtype=np.complex128)
    J = np.zeros((N, N), dtype=np.complex128)
    K = np.zeros((N, N), dtype=np.complex128)
    lib().mdo_jk_incore(N, T.reshape(-1), Pc.view(np.float64).reshape(-1), J.view(np.float64).reshape(-1),
                        K.view(np.float64).reshape(-1))
    return J, K


def onee(bfs, charges, coords, origin):
    """S, T, V (N,N), M (3,N,N), L (3,N,N) as mmd/molecule.py:235-276 assembles them from cython/onee.pyx."""
    fb = _fb(bfs)
    N = fb.nbf
    Z = np.ascontiguousarray(charges, dtype=np.float64)
    xyz = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1)
    org = np.ascontiguousarray(origin, dtype=np.float64)
    S = np.zeros((N, N)); T = np.zeros((N, N)); V = np.zeros((N, N))
    M = np.zeros((3, N, N)); Lm = np.zeros((3, N, N))
    lib().mdo_onee(N, len(Z), Z, xyz, org, S.reshape(-1), T.reshape(-1), V.reshape(-1), M.reshape(-1), Lm.reshape(-1), *fb.args())
    return S, T, V, M, Lm


# ---- nuclear-derivative integrals (cython/grad.pyx; the next row of the scope table) ------------
_CENTER = {"a": 0, "b": 1, "c": 2, "d": 3}


def ERIx_batch(bfs, idx, x, center):
    """d/dX of (ij|kl) for idx[n,4]; x[n] in 0..2, center[n] in 0..3 (or 'a'..'d'): grad.pyx:74-101."""
    fb = _fb(bfs)
    idx = np.ascontiguousarray(idx, dtype=np.int64).reshape(-1, 4)
    cen = [(_CENTER[c.lower()] if isinstance(c, str) else int(c)) for c in np.atleast_1d(center)]
    xc = np.ascontiguousarray(np.stack([np.broadcast_to(np.asarray(x, dtype=np.int64), len(idx)),
                                        np.broadcast_to(np.asarray(cen, dtype=np.int64), len(idx))], axis=1)).reshape(-1)
    out = np.zeros(len(idx))
    lib().mdo_ERIx_batch(fb.nbf, *fb.args(), len(idx), idx.reshape(-1), xc, out)
    return out


def ERIx(a, b, c, d, x=0, center="a"):
    return float(ERIx_batch([a, b, c, d], [[0, 1, 2, 3]], [x], [center])[0])


def Sx(a, b, x=0, center="A"):
    fb = FlatBasis([a, b])
    return lib().mdo_onee_x(2, *fb.args(), 0, 1, int(x), 0 if center.upper() == "A" else 1, 0)


def Tx(a, b, x=0, center="A"):
    fb = FlatBasis([a, b])
    return lib().mdo_onee_x(2, *fb.args(), 0, 1, int(x), 0 if center.upper() == "A" else 1, 1)


def VxA(a, b, C, x=0):
    """Operator (Hellmann-Feynman) derivative of the nuclear attraction integral: grad.pyx:43-54."""
    fb = FlatBasis([a, b])
    return lib().mdo_V_x(2, *fb.args(), 0, 1, np.ascontiguousarray(C, dtype=np.float64), int(x), 2)


def VxB(a, b, C, x=0, center="A"):
    """Basis-function-centre derivative of the nuclear attraction integral: grad.pyx:59-71."""
    fb = FlatBasis([a, b])
    return lib().mdo_V_x(2, *fb.args(), 0, 1, np.ascontiguousarray(C, dtype=np.float64), int(x), 0 if center.upper() == "A" else 1)


def forces(bfs, charges, coords, masks, P, F, only=None):
    """RHF nuclear forces -dE/dX (natom, 3) by the reference's recipe (mmd/forces.py:8-99): derivative one-electron
    matrices, the derivative two-electron tensor with its 8-fold symmetry, 2J-K contraction, energy-weighted
    density for the overlap term and the nuclear repulsion.  NumPy + the C derivative integrals above; the N^4
    tensor per atom and direction makes it a small-molecule checker (as in the reference)."""
    bfs = list(bfs)
    N = len(bfs)
    Z = np.asarray(charges, dtype=np.float64)
    xyz = np.asarray(coords, dtype=np.float64).reshape(-1, 3)
    masks = np.asarray(masks, dtype=np.float64).reshape(len(Z), N)
    P = np.asarray(P)
    F = np.asarray(F)
    fb = FlatBasis(bfs)
    # canonical quartets i>=j, k>=l, ij>=kl (forces.py:61-67)
    i2, j2 = np.tril_indices(N)
    ij = i2 * (i2 + 1) // 2 + j2
    A, B = np.meshgrid(np.arange(len(ij)), np.arange(len(ij)), indexing="ij")
    keep = ij[A] >= ij[B]
    quart = np.stack([i2[A[keep]], j2[A[keep]], i2[B[keep]], j2[B[keep]]], axis=1)
    W = P @ F @ P
    out = np.zeros((len(Z), 3))
    for a in range(len(Z)):
        m = masks[a]
        for x in range(3):
            if only is not None and (a, x) not in only:      # restrict to some (atom, direction) components
                continue
            dS = np.zeros((N, N)); dT = np.zeros((N, N)); dV = np.zeros((N, N))
            for i in range(N):
                for j in range(i + 1):
                    dS[i, j] = dS[j, i] = m[i] * Sx(bfs[i], bfs[j], x, "A") + m[j] * Sx(bfs[i], bfs[j], x, "B")
                    dT[i, j] = dT[j, i] = m[i] * Tx(bfs[i], bfs[j], x, "A") + m[j] * Tx(bfs[i], bfs[j], x, "B")
                    v = -Z[a] * VxA(bfs[i], bfs[j], xyz[a], x)
                    for c in range(len(Z)):
                        v -= m[i] * Z[c] * VxB(bfs[i], bfs[j], xyz[c], x, "A")
                        v -= m[j] * Z[c] * VxB(bfs[i], bfs[j], xyz[c], x, "B")
                    dV[i, j] = dV[j, i] = v
            dVN = 0.0
            for c in range(len(Z)):
                R = np.linalg.norm(xyz[a] - xyz[c])
                if not np.allclose(R, 0.0):
                    dVN += -(xyz[a, x] - xyz[c, x]) * Z[a] * Z[c] / R ** 3
            val = np.zeros(len(quart))
            for cen in range(4):
                w = m[quart[:, cen]]
                sel = np.nonzero(w)[0]
                if len(sel):
                    val[sel] += w[sel] * ERIx_batch(fb, quart[sel], np.full(len(sel), x), np.full(len(sel), cen))
            dG = np.zeros((N, N, N, N))
            i, j, k, l = quart.T
            for p_ in ((i, j, k, l), (k, l, i, j), (j, i, l, k), (l, k, j, i), (j, i, k, l), (l, k, i, j), (i, j, l, k), (k, l, j, i)):
                dG[p_] = val
            Hx = dT + dV
            Jx = np.einsum("pqrs,sr->pq", dG, P)
            Kx = np.einsum("psqr,sr->pq", dG, P)
            Fx = Hx + 2.0 * Jx - Kx
            force = np.einsum("pq,qp", P, Fx + Hx) - 2.0 * np.einsum("pq,qp", dS, W) + dVN
            out[a, x] = np.real(-force)
    return out


def forces_2e_contracted(bfs, masks, P):
    """Two-electron part of dE/dX (natom, 3) WITHOUT the N^4 derivative tensor of mmd/forces.py:61-92: every canonical
    quartet i>=j, k>=l, ij>=kl contributes
        (deg / 8) * [16 P_ij P_kl - 4 P_ik P_jl - 4 P_il P_jk] * sum_{centres c on the atom} d(ij|kl)/dX_c
    with the degeneracy of cython/fock.pyx:60-70.  It equals einsum(P, 2 Jx - Kx) of the reference for a real
    symmetric P (P = C_occ C_occ^T, no factor 2) — the contraction a device gradient kernel performs as it goes
    (SURVEY 8f rank 4); tests/test_oracle_pinned.py checks it against the tensor route."""
    bfs = list(bfs)
    N = len(bfs)
    masks = np.asarray(masks, dtype=np.float64).reshape(-1, N)
    P = np.real(np.asarray(P))
    fb = FlatBasis(bfs)
    i2, j2 = np.tril_indices(N)
    ij = i2 * (i2 + 1) // 2 + j2
    A, B = np.meshgrid(np.arange(len(ij)), np.arange(len(ij)), indexing="ij")
    keep = ij[A] >= ij[B]
    q = np.stack([i2[A[keep]], j2[A[keep]], i2[B[keep]], j2[B[keep]]], axis=1)
    i, j, k, l = q.T
    deg = np.where(i == j, 1.0, 2.0) * np.where(k == l, 1.0, 2.0) * np.where((i == k) & (j == l), 1.0, 2.0)
    w = deg / 8.0 * (16.0 * P[i, j] * P[k, l] - 4.0 * P[i, k] * P[j, l] - 4.0 * P[i, l] * P[j, k])
    out = np.zeros((len(masks), 3))
    for x in range(3):
        for cen in range(4):
            d = ERIx_batch(fb, q, np.full(len(q), x), np.full(len(q), cen))
            out[:, x] += masks[:, q[:, cen]] @ (w * d)
    return out
