"""One-electron integrals with the reference's call signatures (cython/onee.pyx), evaluated on
the device (csrc/onee.cu).  The element-wise functions S/T/V/Mu/RxDel exist for API parity; the
drivers use `one_electron_matrices` which computes every matrix in one launch."""
import numpy as np

from mmd._b200 import engine as _engine

_AX = {"x": 0, "y": 1, "z": 2}


def one_electron_matrices(bfs, charges, coords, origin):
    """-> S, T, V (N,N), M (3,N,N), L (3,N,N)  (mmd/molecule.py:235-276)."""
    return _engine.engine_for(list(bfs)).onee(charges, coords, origin)


def _pair(a, b, charges=(), coords=(), origin=(0.0, 0.0, 0.0)):
    bfs = [a] if a is b else [a, b]
    S_, T_, V_, M_, L_ = _engine.engine_for(bfs).onee(np.asarray(charges, dtype=float),
                                                      np.asarray(coords, dtype=float).reshape(-1, 3), origin)
    if a is b:   # the matrix driver leaves -L_ii on the diagonal (mmd/molecule.py:276); undo for the element call
        return S_[0, 0], T_[0, 0], V_[0, 0], M_[:, 0, 0], -L_[:, 0, 0]
    # element (1,0) is the one evaluated with (bfs[1], bfs[0]); (0,1) is its mirror image
    return S_[0, 1], T_[0, 1], V_[0, 1], M_[:, 0, 1], L_[:, 0, 1]


def S(a, b):
    return float(_pair(a, b)[0])


def T(a, b):
    return float(_pair(a, b)[1])


def V(a, b, C):
    """Nuclear attraction integral for a unit POSITIVE charge at C without the -Z factor, like the
    reference (the caller multiplies by -charge)."""
    return float(-_pair(a, b, charges=[1.0], coords=[np.asarray(C, dtype=float)])[2])


def Mu(a, b, C, direction):
    return float(_pair(a, b, origin=np.asarray(C, dtype=float))[3][_AX[direction.lower()]])


def RxDel(a, b, C, direction):
    return float(_pair(a, b, origin=np.asarray(C, dtype=float))[4][_AX[direction.lower()]])


def _boys(n, T):
    """Boys function F_n(T) as the kernels evaluate it (integer n <= 8)."""
    if int(n) != n or n < 0 or n > 8:
        raise NotImplementedError("device Boys routine covers integer orders 0..8 (the (dd|dd) range)")
    return float(_engine.boys(int(n), [T])[0, int(n)])
