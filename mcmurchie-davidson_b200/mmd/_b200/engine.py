"""Engine: one GPU's two-electron engine for one list of basis functions.

Python owns every user-visible buffer (numpy on the host, torch tensors on the device — torch is
only the device-buffer carrier and the NCCL plumbing); libmmdb200.so owns the shell-pair tables
behind an opaque handle.  All arithmetic happens in the CUDA kernels of csrc/.
"""
import ctypes as C
import collections
import os
import weakref

import numpy as np

from . import dist as D
from . import lib as L
from .shells import ShellTable

_PRIM_CUT = float(os.environ.get("MMDB_PRIM_CUT", "1e-20"))


def _torch():
    import torch
    return torch


class SchwarzTable(dict):
    """The reference's `screen` dict (mmd/molecule.py:95-99): key p(p+1)//2+q -> (pq|pq).
    Remembers which engine already holds it on the device so formPT need not upload it again."""
    engine = None
    flat = None


class Engine(object):
    def __init__(self, bfs, device=None, prim_cut=None):
        L.require_gpu()
        torch = _torch()
        self.lib = L.load()
        self.table = ShellTable(bfs)
        if device is None:
            device = torch.cuda.current_device()
        self.device = int(device)
        self.tdev = torch.device("cuda", self.device)
        t = self.table
        h = C.c_void_p()
        L.check(self.lib.mmdb_basis_create(self.device, t.nshell, L.ptr(t.am), L.ptr(t.nprim), L.ptr(t.poff),
                                           L.ptr(t.centre), L.ptr(t.exps), L.ptr(t.coefs), L.ptr(t.bf0),
                                           _PRIM_CUT if prim_cut is None else float(prim_cut), C.byref(h)))
        self.h = h
        self.N = t.nuser
        self.Ndev = t.ndev
        npairs = np.zeros(L.NCLASS_PAIR, dtype=np.int64)
        nprimpairs = np.zeros(L.NCLASS_PAIR, dtype=np.int64)
        L.check(self.lib.mmdb_basis_pair_counts(self.h, L.ptr(npairs), L.ptr(nprimpairs)))
        self.npairs, self.nprimpairs = npairs, nprimpairs
        # (shell A, shell B) -> (pair class, index in class, flipped?)
        ns = t.nshell
        self.pair_class = np.full((ns, ns), -1, dtype=np.int64)
        self.pair_index = np.full((ns, ns), -1, dtype=np.int64)
        self.pair_flip = np.zeros((ns, ns), dtype=bool)
        for pc in range(L.NCLASS_PAIR):
            n = int(npairs[pc])
            if n == 0:
                continue
            sa = np.zeros(n, dtype=np.int32)
            sb = np.zeros(n, dtype=np.int32)
            L.check(self.lib.mmdb_basis_pair_shells(self.h, pc, L.ptr(sa), L.ptr(sb)))
            idx = np.arange(n)
            self.pair_class[sa, sb] = pc
            self.pair_index[sa, sb] = idx
            self.pair_class[sb, sa] = pc
            self.pair_index[sb, sa] = idx
            self.pair_flip[sb, sa] = True
            self.pair_flip[sa, sb] = False
        self._schwarz = None
        self._resident_screen = None      # the screen object whose values the device currently holds
        self._resident_len = -1
        self.TwoE_dev = None
        self.last_stats = None
        self._pin = {}
        self._keepalive = None
        # MMDB_DETERMINISTIC=1: fixed-point integer accumulation of G (bitwise reproducible builds)
        self.deterministic = os.environ.get("MMDB_DETERMINISTIC", "0") not in ("", "0")

    def __del__(self):
        try:
            if getattr(self, "h", None) is not None and self.h.value:
                self.lib.mmdb_basis_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- helpers -----------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.tdev).cuda_stream)

    def _ncart(self, l):
        return (l + 1) * (l + 2) // 2

    # ---- individual contracted integrals (cython/twoe.pyx:36-50 ERI) -------------------------
    def eri_quartets(self, idx, impl=0):
        """(ij|kl) for user function quartets idx[n,4]."""
        torch = _torch()
        idx = np.asarray(idx, dtype=np.int64).reshape(-1, 4)
        t = self.table
        d = t.user2dev[idx]
        sh = t.fn_shell[d]
        cp = t.fn_comp[d]
        out = np.zeros(len(idx), dtype=np.float64)
        A, B, Cc, D = sh[:, 0], sh[:, 1], sh[:, 2], sh[:, 3]
        pcb, pib, flb = self.pair_class[A, B], self.pair_index[A, B], self.pair_flip[A, B]
        pck, pik, flk = self.pair_class[Cc, D], self.pair_index[Cc, D], self.pair_flip[Cc, D]
        a = np.where(flb, cp[:, 1], cp[:, 0])
        b = np.where(flb, cp[:, 0], cp[:, 1])
        c = np.where(flk, cp[:, 3], cp[:, 2])
        dd = np.where(flk, cp[:, 2], cp[:, 3])
        swap = pcb < pck
        pcb2, pck2 = np.where(swap, pck, pcb), np.where(swap, pcb, pck)
        pib2, pik2 = np.where(swap, pik, pib), np.where(swap, pib, pik)
        a2, b2 = np.where(swap, c, a), np.where(swap, dd, b)
        c2, d2 = np.where(swap, a, c), np.where(swap, b, dd)
        alive = (pcb >= 0) & (pck >= 0)   # pairs dropped by the primitive cut are exactly negligible
        with torch.cuda.device(self.tdev):
            for cb in range(L.NCLASS_PAIR):
                for ck in range(cb + 1):
                    sel = np.nonzero(alive & (pcb2 == cb) & (pck2 == ck))[0]
                    if len(sel) == 0:
                        continue
                    (la, lb), (lc, ld) = L.PAIR_CLASSES[cb], L.PAIR_CLASSES[ck]
                    nb, nc, nd = self._ncart(lb), self._ncart(lc), self._ncart(ld)
                    nfn = self._ncart(la) * nb * nc * nd
                    # unique shell quartets among the requested function quartets
                    key = pib2[sel] * (int(self.npairs[ck]) + 1) + pik2[sel]
                    uk, inv = np.unique(key, return_inverse=True)
                    bi = torch.from_numpy((uk // (int(self.npairs[ck]) + 1)).astype(np.int32)).to(self.tdev)
                    ki = torch.from_numpy((uk % (int(self.npairs[ck]) + 1)).astype(np.int32)).to(self.tdev)
                    buf = torch.empty(len(uk) * nfn, dtype=torch.float64, device=self.tdev)
                    L.check(self.lib.mmdb_eri_shell_quartets(self.h, cb, ck, len(uk), L.ptr(bi), L.ptr(ki), L.ptr(buf),
                                                             int(impl), self._stream()))
                    vals = buf.cpu().numpy().reshape(len(uk), nfn)
                    f = ((a2[sel] * nb + b2[sel]) * nc + c2[sel]) * nd + d2[sel]
                    out[sel] = vals[inv, f]
        return out

    # ---- Schwarz table (mmd/molecule.py:95-99) -------------------------------------------------
    def schwarz(self):
        torch = _torch()
        with torch.cuda.device(self.tdev):
            Q = torch.empty((self.Ndev, self.Ndev), dtype=torch.float64, device=self.tdev)
            L.check(self.lib.mmdb_schwarz(self.h, L.ptr(Q), self._stream()))
            Qh = self.table.to_user_matrix(Q.cpu().numpy())
        N = self.N
        p, q = np.tril_indices(N)
        flat = np.ascontiguousarray(Qh[p, q])        # order p(p+1)/2+q
        tab = SchwarzTable(zip(range(len(flat)), flat.tolist()))
        tab.engine = weakref.ref(self)    # no strong reference: a table must not keep an evicted engine (and its N^4 tensor) alive
        tab.flat = flat
        self._schwarz = tab
        self._resident_screen = tab
        return tab

    def _install_screen(self, screen):
        """Make sure the device holds the caller's `screen` (dict keyed p(p+1)//2+q, or flat array).  The engine
        remembers WHICH table is resident (`_resident_screen`): its own SchwarzTable object, or the identity of a
        caller-supplied one, so a later formPT(..., screen=None) re-installs the engine's own table instead of
        screening with a stale foreign one, and a plain dict is not uploaded again while it stays resident."""
        if screen is None:
            if self._schwarz is None:
                self.schwarz()
                return
            if self._resident_screen is self._schwarz:
                return
            screen = self._schwarz
        if screen is self._resident_screen:
            if not isinstance(screen, dict) or isinstance(screen, SchwarzTable) or len(screen) == self._resident_len:
                return
        N = self.N
        ntri = N * (N + 1) // 2
        if isinstance(screen, SchwarzTable) and screen.flat is not None and len(screen.flat) == ntri:
            flat = np.ascontiguousarray(screen.flat, dtype=np.float64)
        elif isinstance(screen, dict):
            flat = np.fromiter((screen[k] for k in range(ntri)), dtype=np.float64, count=ntri)
        else:
            flat = np.ascontiguousarray(screen, dtype=np.float64)
        if not self.table.identity:
            Qu = np.zeros((N, N))
            p, q = np.tril_indices(N)
            Qu[p, q] = flat
            Qu[q, p] = flat
            Qd = self.table.to_dev_matrix(Qu)
            p, q = np.tril_indices(self.Ndev)
            flat = np.ascontiguousarray(Qd[p, q])
        L.check(self.lib.mmdb_set_schwarz_host(self.h, L.ptr(flat)))
        self._resident_screen = screen           # strong reference: an id() could be recycled by a new object
        self._resident_len = len(screen) if isinstance(screen, dict) else -1

    # ---- dense tensor (cython/twoe.pyx:12-31 doERIs) -------------------------------------------
    def dense(self, keep_device=True):
        torch = _torch()
        n = self.Ndev
        with torch.cuda.device(self.tdev):
            T = torch.empty((n, n, n, n), dtype=torch.float64, device=self.tdev)
            L.check(self.lib.mmdb_eri_dense(self.h, L.ptr(T), self._stream()))
            host = T.cpu().numpy()
        if self.table.identity:
            if keep_device:
                self.TwoE_dev = T
            return host
        host = self.table.to_user_tensor4(host)
        if keep_device:
            self.TwoE_dev = torch.from_numpy(host).to(self.tdev)
        return host

    def dense_device(self):
        """Fill the dense tensor on the device only (self.TwoE_dev); False when the caller's function order differs from
        the device order (then dense() with its host-side reordering is the way)."""
        if not self.table.identity:
            return False
        torch = _torch()
        n = self.Ndev
        with torch.cuda.device(self.tdev):
            T = torch.empty((n, n, n, n), dtype=torch.float64, device=self.tdev)
            L.check(self.lib.mmdb_eri_dense(self.h, L.ptr(T), self._stream()))
        self.TwoE_dev = T
        return True

    def dense_host(self):
        """Host copy of the resident tensor (rebuilt when the engine has none)."""
        if self.TwoE_dev is None or not self.table.identity:
            return self.dense(keep_device=True)
        return self.TwoE_dev.cpu().numpy()

    # ---- in-core J/K (mmd/scf.py:97-98) ----------------------------------------------------------
    def jk_incore(self, P, TwoE=None):
        """J, K (complex128 (N,N)) from the dense tensor; P complex or real (N,N), user order."""
        torch = _torch()
        N = self.N
        with torch.cuda.device(self.tdev):
            if TwoE is not None:
                T = TwoE if not isinstance(TwoE, np.ndarray) else torch.from_numpy(np.ascontiguousarray(TwoE)).to(self.tdev)
            else:
                if self.TwoE_dev is None:
                    self.dense(keep_device=True)          # rebuilt rather than failing (e.g. a fresh engine after a cache eviction)
                T = self.TwoE_dev
            P = np.asarray(P)
            Pre = torch.from_numpy(np.ascontiguousarray(P.real, dtype=np.float64)).to(self.tdev)
            cplx = np.iscomplexobj(P) and bool(np.any(P.imag != 0.0))
            Pim = torch.from_numpy(np.ascontiguousarray(P.imag, dtype=np.float64)).to(self.tdev) if cplx else None
            out = torch.empty((4, N, N), dtype=torch.float64, device=self.tdev)
            L.check(self.lib.mmdb_jk_incore(self.device, L.ptr(T), N, L.ptr(Pre), L.ptr(Pim), L.ptr(out[0]),
                                            L.ptr(out[1]) if cplx else None, L.ptr(out[2]),
                                            L.ptr(out[3]) if cplx else None, self._stream()))
            o = out.cpu().numpy()
        J = o[0].astype(np.complex128)
        K = o[2].astype(np.complex128)
        if cplx:
            J += 1j * o[1]
            K += 1j * o[3]
        return J, K

    # ---- AO->MO transformation + MP2 (mmd/postscf.py:21-41, 59-70) ---------------------------------
    def ao2mo_mp2(self, Cmat, eps, nocc, want_energy=True):
        """single_bar (N,N,N,N) in the MO basis and the MP2 correlation energy, from the device-resident
        dense tensor.  Real orbital coefficients only (the caller keeps the host path otherwise)."""
        torch = _torch()
        if self.TwoE_dev is None:
            raise L.MMDBError("ao2mo_mp2: no device-resident TwoE (run the in-core path first)")
        N = self.N
        with torch.cuda.device(self.tdev):
            Cd = torch.from_numpy(np.ascontiguousarray(Cmat, dtype=np.float64)).to(self.tdev)
            ed = torch.from_numpy(np.ascontiguousarray(eps, dtype=np.float64)).to(self.tdev)
            MO = torch.empty((N, N, N, N), dtype=torch.float64, device=self.tdev)
            work = torch.empty((N, N, N, N), dtype=torch.float64, device=self.tdev)
            e2 = C.c_double(0.0)
            L.check(self.lib.mmdb_ao2mo_mp2(self.device, L.ptr(self.TwoE_dev), N, int(nocc), L.ptr(Cd), L.ptr(ed), L.ptr(MO),
                                            L.ptr(work), C.byref(e2) if want_energy else None, self._stream()))
            del work
            single_bar = MO.cpu().numpy()
        return single_bar, (e2.value if want_energy else None)

    # ---- direct Fock build (cython/fock.pyx:13-87 formPT) ----------------------------------------
    def _pinned(self, name, shape, dtype=float):
        """Cached page-locked staging buffer (torch tensor + numpy view)."""
        torch = _torch()
        cur = self._pin.get(name)
        tdt = torch.complex128 if dtype is complex else torch.float64
        if cur is None or tuple(cur[0].shape) != tuple(shape) or cur[0].dtype != tdt:
            t = torch.empty(shape, dtype=tdt, pin_memory=True)
            cur = (t, t.numpy())
            self._pin[name] = cur
        return cur

    def _fock_direct_planes(self, dP, cplx, tol, want_stats, flags):
        """dP: device float64 (nplane, n, n) in device order -> un-symmetrised G planes (device), summed
        over the ranks (one all-reduce over NCCL / NVLink: mmdb_allreduce_G of the C ABI on GPUs; the
        torch.distributed collective only where there is no NCCL, i.e. the gloo CPU tests)."""
        torch = _torch()
        n = self.Ndev
        rank, world = D.world()
        nplane = 2 if cplx else 1
        G = torch.zeros((nplane, n, n), dtype=torch.float64, device=self.tdev)
        stats = L.FockStats() if want_stats else None
        if self.deterministic:
            flags = int(flags) | 2
        L.check(self.lib.mmdb_fock_direct(self.h, L.ptr(dP[0]), L.ptr(dP[1]) if cplx else None, float(tol), L.ptr(G[0]),
                                          L.ptr(G[1]) if cplx else None, rank, world, int(flags),
                                          C.byref(stats) if want_stats else None, self._stream()))
        if world > 1:
            if self.tdev.type == "cuda" and not os.environ.get("MMDB_TORCH_ALLREDUCE"):
                comm = D.c_abi_comm(self.lib, self.device)
                L.check(self.lib.mmdb_allreduce_G(comm, L.ptr(G), G.numel(), 1 if self.deterministic else 0, self._stream()))
            else:
                D.allreduce_sum_(G.view(torch.int64) if self.deterministic else G)
        if self.deterministic:
            L.check(self.lib.mmdb_fixed_to_double(self.device, L.ptr(G), G.numel(), self._stream()))
        self._last_stats_raw = stats
        return G

    def formPT(self, P, P_old, screen=None, tol=1e-12, want_stats=True, flags=0):
        """Un-symmetrised G (complex128, user order).  Shards over torch.distributed ranks when a
        process group with world_size > 1 is initialised (one process per GPU) and sums the partial
        G matrices with one all-reduce (NCCL over NVLink).

        Host side: ONE pass over the inputs (dP = P - P_old, fock.pyx:24, split into page-locked real / imaginary
        planes by mmdb_c128_diff_split_host) and one pass that interleaves the page-locked G planes into the result.
        A real density moves one plane each way (8 N^2 bytes H2D, 8 N^2 bytes D2H); nothing synchronises before the
        final read-back."""
        torch = _torch()
        self._install_screen(screen)
        P = np.asarray(P)
        P_old = np.asarray(P_old)
        n = self.Ndev
        pin_in, pin_in_np = self._pinned("dP", (2, n, n))
        pin_out, pin_out_np = self._pinned("G", (2, n, n))
        if not self.table.identity:
            P, P_old = self.table.to_dev_matrix(P), self.table.to_dev_matrix(P_old)
        P = np.ascontiguousarray(P, dtype=np.complex128)
        P_old = np.ascontiguousarray(P_old, dtype=np.complex128)
        has_im = C.c_int(0)
        L.check(self.lib.mmdb_c128_diff_split_host(P.ctypes.data, P_old.ctypes.data, n * n, pin_in_np[0].ctypes.data,
                                                   pin_in_np[1].ctypes.data, C.byref(has_im)))
        cplx = bool(has_im.value)
        npl = 2 if cplx else 1
        with torch.cuda.device(self.tdev):
            dP = pin_in[:npl].to(self.tdev, non_blocking=True)
            G = self._fock_direct_planes(dP, cplx, tol, want_stats, flags)
            pin_out[:npl].copy_(G, non_blocking=True)
            torch.cuda.current_stream(self.tdev).synchronize()
        self.last_stats = self._last_stats_raw.as_dict() if want_stats else None
        out = np.empty((n, n), dtype=np.complex128)
        L.check(self.lib.mmdb_c128_join_host(pin_out_np[0].ctypes.data, pin_out_np[1].ctypes.data if cplx else None, n * n,
                                             out.ctypes.data))
        self.h2d_bytes = self.d2h_bytes = 8 * n * n * npl          # what this call moved over PCIe (bench.py reports it)
        return self.table.to_user_matrix(out)

    # ---- device-resident variants for the SCF driver: no host round trip ------
    @property
    def supports_device_scf(self):
        """Device and user function order coincide (always true for Molecule-built basis lists)."""
        return bool(self.table.identity)

    def formPT_dev(self, P, P_old, screen=None, tol=1e-12, want_stats=False, flags=0):
        """formPT on device tensors: P, P_old complex128 (N,N) on self.tdev -> un-symmetrised complex128 G."""
        torch = _torch()
        if not self.table.identity:
            raise L.MMDBError("formPT_dev: function order differs between user and device")
        self._install_screen(screen)
        with torch.cuda.device(self.tdev):
            d = P - P_old
            cplx = bool(d.is_complex() and torch.any(d.imag != 0.0).item())
            if cplx:
                dP = torch.stack((d.real, d.imag)).contiguous()
            else:
                dP = (d.real if d.is_complex() else d).contiguous().unsqueeze(0)
            G = self._fock_direct_planes(dP, cplx, tol, want_stats, flags)
            self.last_stats = self._last_stats_raw.as_dict() if want_stats else None
            return torch.complex(G[0], G[1] if cplx else torch.zeros_like(G[0]))

    def jk_incore_dev(self, P):
        """J, K complex128 (N,N) device tensors from the device-resident dense tensor; P complex128 on device."""
        torch = _torch()
        if self.TwoE_dev is None:
            raise L.MMDBError("jk_incore_dev: no device-resident TwoE (call dense() first)")
        N = self.N
        with torch.cuda.device(self.tdev):
            Pre = (P.real if P.is_complex() else P).contiguous()
            cplx = bool(P.is_complex() and torch.any(P.imag != 0.0).item())
            Pim = P.imag.contiguous() if cplx else None
            out = torch.zeros((4, N, N), dtype=torch.float64, device=self.tdev)
            L.check(self.lib.mmdb_jk_incore(self.device, L.ptr(self.TwoE_dev), N, L.ptr(Pre), L.ptr(Pim), L.ptr(out[0]),
                                            L.ptr(out[1]) if cplx else None, L.ptr(out[2]),
                                            L.ptr(out[3]) if cplx else None, self._stream()))
            return torch.complex(out[0], out[1]), torch.complex(out[2], out[3])

    # ---- one-electron integrals (cython/onee.pyx) ------------------------------------------------
    def onee(self, charges, coords, origin):
        Z = np.ascontiguousarray(charges, dtype=np.float64)
        xyz = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1)
        org = np.ascontiguousarray(origin, dtype=np.float64)
        n = self.Ndev
        S = np.zeros((n, n)); T = np.zeros((n, n)); V = np.zeros((n, n))
        M = np.zeros((3, n, n)); Lm = np.zeros((3, n, n))
        L.check(self.lib.mmdb_onee_host(self.h, len(Z), L.ptr(Z), L.ptr(xyz), L.ptr(org), L.ptr(S), L.ptr(T), L.ptr(V),
                                        L.ptr(M), L.ptr(Lm)))
        if self.table.identity:
            return S, T, V, M, Lm
        u = self.table.user2dev
        ix = np.ix_(u, u)
        return (np.ascontiguousarray(S[ix]), np.ascontiguousarray(T[ix]), np.ascontiguousarray(V[ix]),
                np.ascontiguousarray(M[:, u][:, :, u]), np.ascontiguousarray(Lm[:, u][:, :, u]))


    # ---- nuclear gradient (cython/grad.pyx + mmd/forces.py) ----------------------------------------
    def gradient(self, charges, coords, atom_of_function, P, F):
        """dE/dX (natom, 3) of the RHF energy split into (one-electron, two-electron, nuclear repulsion) parts.
        P = C_occ C_occ^T and F in user function order; atom_of_function[i] = atom index of basis function i."""
        Z = np.ascontiguousarray(charges, dtype=np.float64)
        xyz = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1)
        natom = len(Z)
        t = self.table
        if self._schwarz is None:
            self.schwarz()
        P = np.real(np.asarray(P))
        W = P @ np.real(np.asarray(F)) @ P                      # energy-weighted density (mmd/forces.py:94)
        Pd = np.ascontiguousarray(t.to_dev_matrix(np.ascontiguousarray(P)), dtype=np.float64)
        Wd = np.ascontiguousarray(t.to_dev_matrix(np.ascontiguousarray(W)), dtype=np.float64)
        atom_of_function = np.asarray(atom_of_function)
        shell_atom = np.zeros(t.nshell, dtype=np.int32)
        for s in range(t.nshell):
            users = t.dev2user[t.bf0[s]:t.bf0[s] + (int(t.am[s]) + 1) * (int(t.am[s]) + 2) // 2]
            users = users[users >= 0]
            shell_atom[s] = int(atom_of_function[users[0]])
        g1 = np.zeros((natom, 3)); g2 = np.zeros((natom, 3)); gn = np.zeros((natom, 3))
        L.check(self.lib.mmdb_gradient_host(self.h, natom, L.ptr(Z), L.ptr(xyz), L.ptr(shell_atom), L.ptr(Pd), L.ptr(Wd),
                                            L.ptr(g1), L.ptr(g2), L.ptr(gn)))
        return g1, g2, gn


# ---- engines -------------------------------------------------------------------------------------------------
# A Molecule OWNS its engine (mmd/molecule.py keeps it in self._engine, rebuilt by formBasis), so nothing another
# molecule or an element-wise call does can evict it.  Everything else (the module-level ERI / S / T / V / formPT /
# doERIs functions called with an arbitrary list of Basis objects) goes through a small LRU cache keyed by the identity
# of the Basis objects; one entry is evicted at a time, and engines registered by a live Molecule are found first.
_CACHE = collections.OrderedDict()
_CACHE_MAX = 8
_OWNED = weakref.WeakValueDictionary()      # key -> engine owned by a live Molecule


def _key(bfs):
    return tuple(id(b) for b in bfs)


def register_owned(bfs, eng):
    eng._keepalive = list(bfs)
    _OWNED[_key(bfs)] = eng


def owned_engine(bfs):
    """A fresh engine for a Molecule to own (never shared through the LRU cache)."""
    eng = Engine(bfs)
    register_owned(bfs, eng)
    return eng


def engine_for(bfs):
    key = _key(bfs)
    hit = _OWNED.get(key)
    if hit is not None:
        return hit
    hit = _CACHE.get(key)
    if hit is not None:
        _CACHE.move_to_end(key)
        return hit
    eng = Engine(bfs)
    eng._keepalive = list(bfs)
    _CACHE[key] = eng
    while len(_CACHE) > _CACHE_MAX:
        _CACHE.popitem(last=False)           # least recently used only
    return eng


def engine_containing(fns):
    """An existing engine whose basis list contains every function of `fns` (by identity) and their positions in
    it — lets ERI(a,b,c,d) / S / T / V on functions of a molecule reuse the molecule's engine instead of
    building a new one per distinct tuple (a loop like the reference's mmd/molecule.py:95-99 would otherwise
    create thousands of engines)."""
    ids = [id(f) for f in fns]
    for eng in list(_OWNED.values()) + list(reversed(_CACHE.values())):
        pos = getattr(eng, "_pos_of", None)
        if pos is None:
            pos = eng._pos_of = {id(b): k for k, b in enumerate(eng._keepalive)}
        if all(i in pos for i in ids):
            return eng, [pos[i] for i in ids]
    return None, None


def boys(n, T):
    """F_n(T) from the device Boys routine (test hook, mirrors onee._boys of the reference)."""
    L.require_gpu()
    Ts = np.ascontiguousarray(np.atleast_1d(np.asarray(T, dtype=np.float64)))
    n = int(n)
    out = np.zeros((len(Ts), n + 1))
    L.check(L.load().mmdb_boys_host(_torch().cuda.current_device(), n, len(Ts), L.ptr(Ts), L.ptr(out)))
    return out


def boys_class(L_tot, T):
    """F_0..F_L(T) through the class kernels' templated Boys routine for total angular momentum L_tot
    (its own table/asymptotic switch-over T_max(L)); test hook."""
    L.require_gpu()
    Ts = np.ascontiguousarray(np.atleast_1d(np.asarray(T, dtype=np.float64)))
    n = int(L_tot)
    out = np.zeros((len(Ts), n + 1))
    L.check(L.load().mmdb_boys_class_host(_torch().cuda.current_device(), n, len(Ts), L.ptr(Ts), L.ptr(out)))
    return out


def fp64_peak(device=0):
    L.require_gpu()
    tf = C.c_double(0.0)
    ms = C.c_float(0.0)
    L.check(L.load().mmdb_fp64_peak(int(device), C.byref(tf), C.byref(ms)))
    return tf.value, ms.value
