"""Post-SCF: AO->MO transformation and MP2 on top of a converged in-core RHF (consumers of the
dense mol.TwoE tensor the B200 engine produced).  Mirrors PostSCF(mol).MP2() of the reference's
mmd/postscf.py:15-72.  Determinant-based methods (CIS/TDHF/CISD/FCI) are outside the hot path this
package accelerates and are not provided."""
import sys
from itertools import product

import numpy as np


EAGER_DOUBLE_BAR_BYTES = 512 * 1024 * 1024       # N <= 38: mol.double_bar is built in ao2mo like the reference does


class PostSCF(object):
    def __init__(self, mol):
        self.mol = mol
        if not self.mol.is_converged:
            sys.exit("SCF not converged, skipping Post-SCF")
        # (the flag first: reading mol.TwoE would copy a device-resident tensor to the host just to see that it exists)
        if not (getattr(self.mol, "_TwoE_on_device", False) or hasattr(self.mol, "TwoE")):
            sys.exit("Post-SCF needs the in-core tensor: run RHF(direct=False)")
        self.ao2mo()

    def ao2mo(self):
        """(pq|rs) -> MO basis by four quarter transformations; mol.single_bar[P,Q,R,S].

        Bra orbitals (first and third index of the chemists' (PQ|RS)) are complex-conjugated.  For real
        orbitals this is the reference's transformation (mmd/postscf.py:26-27); when LAPACK returns a
        complex rotation inside a degenerate orbital set (X = S^-1/2 carries ~1e-17 imaginary noise,
        e.g. CH4) it keeps E(MP2) invariant, where the un-conjugated form would not be."""
        C = self.mol.C
        self._e2_device = None
        eng = getattr(self.mol, "engine", None)
        real_orbitals = not np.iscomplexobj(C) or float(np.abs(C.imag).max()) < 1e-13
        if real_orbitals and eng is not None and getattr(eng, "TwoE_dev", None) is not None and hasattr(eng, "ao2mo_mp2"):
            # device path: four cuBLAS DGEMM quarter transformations on the resident tensor + MP2 reduction kernel
            self.mol.single_bar, self._e2_device = eng.ao2mo_mp2(np.real(C), np.real(self.mol.MO), self.mol.nocc)
        else:
            Cc = np.conjugate(C)
            t = np.einsum("pqrs,sS->pqrS", self.mol.TwoE, C, optimize=True)
            t = np.einsum("pqrS,rR->pqRS", t, Cc, optimize=True)
            t = np.einsum("pqRS,qQ->pQRS", t, C, optimize=True)
            self.mol.single_bar = np.einsum("pQRS,pP->PQRS", t, Cc, optimize=True)
        self.mol.norb = self.mol.nbasis * 2
        self._spin = np.eye(2)
        # The reference sets mol.double_bar right here (mmd/postscf.py:34-35) — code that reads it after PostSCF(mol) keeps
        # working for the sizes the reference itself can handle; beyond that the 16 (2N)^4-byte spin-orbital tensor
        # is built on first use instead (spin_orbital=True), since the spatial-orbital MP2 never needs it.
        self.mol.double_bar = None
        if 16.0 * (2.0 * self.mol.nbasis) ** 4 <= EAGER_DOUBLE_BAR_BYTES:
            self._ensure_double_bar()
        self.mol.fs = np.kron(np.diag(self.mol.MO), self._spin)
        self.mol.Hp = np.kron(np.einsum("uj,vi,uv", C, C, self.mol.Core).real, self._spin)

    def _ensure_double_bar(self):
        if getattr(self.mol, "double_bar", None) is None:
            block = np.kron(np.kron(self.mol.single_bar, self._spin).transpose(), self._spin).real
            self.mol.double_bar = block.transpose(0, 2, 1, 3) - block.transpose(0, 2, 3, 1)

    def MP2(self, spin_orbital=False):
        mol = self.mol
        if spin_orbital:
            self._ensure_double_bar()
            acc = 0.0
            occ, virt = range(mol.nelec), range(mol.nelec, mol.norb)
            for i, j, a, b in product(occ, occ, virt, virt):
                acc += mol.double_bar[i, j, a, b] ** 2 / (mol.fs[i, i] + mol.fs[j, j] - mol.fs[a, a] - mol.fs[b, b])
            mol.emp2 = 0.25 * acc + mol.energy
        elif getattr(self, "_e2_device", None) is not None:
            mol.emp2 = self._e2_device + mol.energy
        else:
            acc = 0.0
            occ, virt = range(mol.nocc), range(mol.nocc, mol.nbasis)
            g = mol.single_bar
            for i, j, a, b in product(occ, occ, virt, virt):
                denom = mol.MO[i] + mol.MO[j] - mol.MO[a] - mol.MO[b]
                acc += np.conjugate(g[i, a, j, b]) * (2.0 * g[i, a, j, b] - g[i, b, j, a]) / denom
            mol.emp2 = acc + mol.energy
        print("E(MP2) = ", mol.emp2.real)
