"""Test-only stand-in for mmd._b200.engine.Engine backed by the CPU oracle.

Lets the CPU test-suite exercise the HOST logic of the drop-in package (Molecule / SCF / PostSCF
drivers, shell tables, sharding) without a GPU.  It lives under tests/ on purpose: the product never
imports the oracle, and these tests say nothing about the CUDA path (the `-m gpu` tests do that).
"""
import numpy as np

from oracle import oracle as O


class OracleEngine(object):
    def __init__(self, bfs):
        self.bfs = list(bfs)
        self.N = len(self.bfs)
        self.fb = O.FlatBasis(self.bfs)
        self.TwoE = None
        self.last_stats = None

    def onee(self, charges, coords, origin):
        return O.onee(self.fb, charges, coords, origin)

    def schwarz(self):
        flat = O.schwarz(self.fb)
        return dict(zip(range(len(flat)), flat.tolist()))

    def dense(self, keep_device=True):
        T = np.zeros((self.N,) * 4)
        O.doERIs(self.N, T, self.fb)
        self.TwoE = T
        return T

    def jk_incore(self, P, TwoE=None):
        return O.jk_incore(self.TwoE if TwoE is None else TwoE, np.asarray(P, dtype=np.complex128))

    def formPT(self, P, P_old, screen=None, tol=1e-12, **kw):
        return O.formPT(P, P_old, self.fb, self.N, screen, tol)

    def eri_quartets(self, idx, impl=0):
        return O.ERI_batch(self.fb, idx)


class OracleTensorEngine(OracleEngine):
    """OracleEngine that also offers the device-tensor entry points of Engine (formPT_dev, jk_incore_dev) on
    CPU torch tensors, so the CPU suite can drive the device-resident SCF loop (mmd/scf.py:_RHF_device)."""
    supports_device_scf = True

    def __init__(self, bfs):
        import torch
        OracleEngine.__init__(self, bfs)
        self.tdev = torch.device("cpu")

    def formPT_dev(self, P, P_old, screen=None, tol=1e-12, want_stats=False, flags=0):
        import torch
        G = O.formPT(P.numpy(), P_old.numpy(), self.fb, self.N, screen, tol)
        return torch.from_numpy(np.ascontiguousarray(G))

    def jk_incore_dev(self, P):
        import torch
        J, K = O.jk_incore(self.TwoE, np.ascontiguousarray(P.numpy()))
        return torch.from_numpy(np.ascontiguousarray(J)), torch.from_numpy(np.ascontiguousarray(K))


def install(monkeypatch, engine_cls=None):
    """Route engine_for() of the drop-in package to the oracle for the duration of a test."""
    cache = {}

    def engine_for(bfs):
        key = tuple(id(b) for b in bfs)
        if key not in cache:
            cache[key] = (engine_cls or OracleEngine)(bfs)
            cache[key]._keep = list(bfs)
        return cache[key]

    import mmd._b200.engine as eng
    import mmd.integrals.fock as fock
    import mmd.integrals.onee as onee
    import mmd.integrals.twoe as twoe
    import mmd.molecule as molecule
    monkeypatch.setattr(eng, "engine_for", engine_for)
    monkeypatch.setattr(molecule, "engine_for", engine_for)
    for mod in (fock, onee, twoe):
        monkeypatch.setattr(mod._engine, "engine_for", engine_for)
    return engine_for
