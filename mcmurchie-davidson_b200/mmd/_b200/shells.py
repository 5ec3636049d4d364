"""Rebuild shell structure from a list of reference-style Basis objects.

The reference works on individual contracted Cartesian functions (cython/basis.pxi); consecutive
Basis objects sharing origin/exponents/contraction are one shell (mmd/molecule.py:64-72 is the only
place shell structure exists).  The device engine works on shells, so this module groups the list
back into shells and records the map between the caller's function order ("user index") and the
device function order (shell by shell, components in the reference's momentum2shell order).

A Basis that is not part of a complete, in-order shell (hand-built lists as in the reference's
tests/test013.py) becomes its own shell whose other Cartesian components are "ghost" functions.
"""
import math

import numpy as np

MAX_AM = 2


def cart_components(l):
    """Cartesian exponents of a shell in the reference's order (mmd/molecule.py:108-114)."""
    return [(i, j, l - i - j) for i in range(l, -1, -1) for j in range(l - i, -1, -1)]


def _fact2(n):
    r = 1
    while n > 1:
        r *= n
        n -= 2
    return r


def comp_factor(lmn):
    """sqrt((2l-1)!!(2m-1)!!(2n-1)!!): removes the per-component part of the primitive norm."""
    l, m, n = (int(x) for x in lmn)
    return math.sqrt(_fact2(2 * l - 1) * _fact2(2 * m - 1) * _fact2(2 * n - 1))


class ShellTable(object):
    def __init__(self, bfs):
        self.nuser = len(bfs)
        am, nprim, poff, centre, exps, coefs, bf0 = [], [], [], [], [], [], []
        dev2user = []
        f = 0
        nprim_total = 0
        while f < self.nuser:
            b = bfs[f]
            lmn = tuple(int(x) for x in b.shell)
            L = sum(lmn)
            if L > MAX_AM:
                raise NotImplementedError("angular momentum > d is not supported by the B200 two-electron engine")
            comps = cart_components(L)
            full = self._is_full_shell(bfs, f, comps)
            e = np.asarray(b.exps, dtype=np.float64)
            c = np.asarray(b.norm, dtype=np.float64) * np.asarray(b.coefs, dtype=np.float64) * comp_factor(lmn)
            am.append(L)
            nprim.append(len(e))
            poff.append(nprim_total)
            nprim_total += len(e)
            centre.append(np.asarray(b.origin, dtype=np.float64))
            exps.append(e)
            coefs.append(c)
            bf0.append(len(dev2user))
            if full:
                dev2user.extend(range(f, f + len(comps)))
                f += len(comps)
            else:
                me = comps.index(lmn)
                dev2user.extend([f if k == me else -1 for k in range(len(comps))])
                f += 1
        self.nshell = len(am)
        self.am = np.asarray(am, dtype=np.int32)
        self.nprim = np.asarray(nprim, dtype=np.int32)
        self.poff = np.asarray(poff, dtype=np.int32)
        self.centre = np.ascontiguousarray(np.concatenate(centre), dtype=np.float64)
        self.exps = np.ascontiguousarray(np.concatenate(exps), dtype=np.float64)
        self.coefs = np.ascontiguousarray(np.concatenate(coefs), dtype=np.float64)
        self.bf0 = np.asarray(bf0, dtype=np.int32)
        self.dev2user = np.asarray(dev2user, dtype=np.int64)
        self.ndev = len(dev2user)
        self.user2dev = np.full(self.nuser, -1, dtype=np.int64)
        real = self.dev2user >= 0
        self.user2dev[self.dev2user[real]] = np.nonzero(real)[0]
        self.identity = bool(self.ndev == self.nuser and np.array_equal(self.dev2user, np.arange(self.nuser)))
        # device function -> (shell, component)
        self.fn_shell = np.repeat(np.arange(self.nshell), [(l + 1) * (l + 2) // 2 for l in am]).astype(np.int64)
        self.fn_comp = (np.arange(self.ndev) - self.bf0[self.fn_shell]).astype(np.int64)

    @staticmethod
    def _is_full_shell(bfs, f, comps):
        if f + len(comps) > len(bfs):
            return False
        b0 = bfs[f]
        raw0 = getattr(b0, "_raw_coefs", None)
        for k, lmn in enumerate(comps):
            b = bfs[f + k]
            if tuple(int(x) for x in b.shell) != lmn:
                return False
            if k == 0:
                continue
            if int(b.num_exps) != int(b0.num_exps):
                return False
            if not (np.array_equal(np.asarray(b.origin), np.asarray(b0.origin)) and
                    np.array_equal(np.asarray(b.exps), np.asarray(b0.exps))):
                return False
            raw = getattr(b, "_raw_coefs", None)
            if raw0 is not None and raw is not None:
                if not np.array_equal(raw, raw0):
                    return False
            else:
                # no raw coefficients kept: compare the component-independent product norm*coef*factor
                x0 = np.asarray(b0.norm) * np.asarray(b0.coefs) * comp_factor(b0.shell)
                x = np.asarray(b.norm) * np.asarray(b.coefs) * comp_factor(b.shell)
                if not np.allclose(x, x0, rtol=1e-13, atol=0.0):
                    return False
        return True

    # ---- matrix / tensor maps between user and device function order -------------------------
    def to_dev_matrix(self, M):
        if self.identity:
            return np.ascontiguousarray(M)
        out = np.zeros((self.ndev, self.ndev), dtype=M.dtype)
        u = self.user2dev
        out[np.ix_(u, u)] = M
        return out

    def to_user_matrix(self, M):
        if self.identity:
            return M
        u = self.user2dev
        return np.ascontiguousarray(M[np.ix_(u, u)])

    def to_user_tensor4(self, T):
        if self.identity:
            return T
        u = self.user2dev
        return np.ascontiguousarray(T[np.ix_(u, u, u, u)])
