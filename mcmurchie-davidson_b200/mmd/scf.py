"""Closed-shell SCF driver.  Behaviour (iteration structure, DIIS, convergence test, printed
summary, attribute names) follows the reference's mmd/scf.py so energies and iteration counts can
be compared mode-for-mode; the two Fock-build branches call the B200 engine:

    direct=True   G = formPT(P, P_old, ...)   fused screened ERI + J/K digestion kernels
    direct=False  J, K from one pass over the device-resident dense tensor (mmd/scf.py:97-98)
"""
import numpy as np
import scipy.linalg
from numpy.linalg import multi_dot

from mmd.integrals.fock import formPT

MAX_SCF_ITER = 100
DIIS_DEPTH = 8


class SCF(object):
    def RHF(self, doPrint=True, DIIS=True, direct=False, conver=1e-8, acc2e=1e-12):
        """Restricted Hartree-Fock for a closed-shell molecule."""
        n = self.nbasis
        self.is_converged = False
        self.delta_energy = 1e20
        self.P_RMS = 1e20
        self.P_old = np.zeros((n, n), dtype="complex")
        self.maxiter = MAX_SCF_ITER
        self.direct = direct
        if self.direct:
            self.incFockRst = False          # True would rebuild G from the full density every step
        self.scrTol = acc2e
        self.build(self.direct)

        self.P = self.P_old
        self.F = self.Core.astype("complex")
        if DIIS:
            self.fockSet, self.errorSet = [], []
        self.scf_history = []                 # (energy, P_RMS) per step — parity diagnostics

        for step in range(self.maxiter):
            if step > 0:
                self.F_old = self.F
                energy_old = self.energy
                self.buildFock()              # uses P and (incremental mode) P_old
                self.P_old = self.P
                if DIIS:
                    # the extrapolated Fock matrix only produces the next density; self.F stays
                    # the un-extrapolated one so the incremental build remains consistent
                    F_diis = self.updateDIIS(self.F, self.P)
                    self.FO = np.dot(self.X.T, np.dot(F_diis, self.X))
            if not DIIS or step == 0:
                self.orthoFock()

            eps, self.CO = scipy.linalg.eigh(self.FO)
            C = np.dot(self.X, self.CO)
            self.C = np.dot(self.X, self.CO)
            self.MO = eps
            occ = C[:, :self.nocc]
            self.P = np.dot(occ, np.conjugate(occ).T)
            self.computeEnergy()

            if step > 0:
                self.delta_energy = self.energy - energy_old
                self.P_RMS = np.linalg.norm(self.P - self.P_old)
            self.scf_history.append((complex(self.energy).real, float(np.real(self.P_RMS))))
            last = step == (self.maxiter - 1)
            if np.abs(self.P_RMS) < conver or last:
                if last:
                    print("NOT CONVERGED")
                    break
                self.is_converged = True
                FPS = np.dot(self.F, np.dot(self.P, self.S))
                residual = FPS - self.adj(FPS)
                self.computeDipole()
                if doPrint:
                    print("E(SCF)    = ", "{0:.12f}".format(self.energy.real) + " in " + str(step) + " iterations")
                    print("  Convergence:")
                    print("    FPS-SPF  = ", np.linalg.norm(residual))
                    print("    RMS(P)   = ", "{0:.2e}".format(self.P_RMS.real))
                    print("    dE(SCF)  = ", "{0:.2e}".format(self.delta_energy.real))
                    print("  Dipole X = ", "{0:.8f}".format(self.mu[0].real))
                    print("  Dipole Y = ", "{0:.8f}".format(self.mu[1].real))
                    print("  Dipole Z = ", "{0:.8f}".format(self.mu[2].real))
                self.scf_iterations = step
                break

    # ---- Fock builds -------------------------------------------------------------------------
    def buildFock(self):
        core = self.Core.astype("complex")
        if self.direct:
            restart = self.incFockRst
            P_ref = np.zeros_like(self.P) if restart else self.P_old
            G = formPT(self.P, P_ref, self.bfs, self.nbasis, self.screen, self.scrTol)
            self.G = 0.5 * (G + G.T)                      # plain transpose, as in the reference
            self.F = (core if restart else self.F_old) + self.G
        else:
            self.J, self.K = self.engine.jk_incore(self.P)
            self.G = 2.0 * self.J - self.K
            self.F = core + self.G

    def orthoFock(self):
        self.FO = np.dot(self.X.T, np.dot(self.F, self.X))

    def unOrthoFock(self):
        self.F = np.dot(self.U.T, np.dot(self.FO, self.U))

    def orthoDen(self):
        self.PO = np.dot(self.U, np.dot(self.P, self.U.T))

    def unOrthoDen(self):
        self.P = np.dot(self.X, np.dot(self.PO, self.X.T))

    def updateFock(self):
        """Rebuild F from the orthonormal-basis density PO (used by propagators / external fields)."""
        self.unOrthoDen()
        self.buildFock()
        self.orthoFock()

    # ---- observables -------------------------------------------------------------------------
    def computeEnergy(self):
        self.el_energy = np.einsum("pq,qp", self.Core + self.F, self.P)
        self.energy = self.el_energy + self.nuc_energy

    def computeDipole(self):
        self.el_energy = np.einsum("pq,qp", self.Core + self.F, self.P)
        for k in range(3):
            nuclear = sum(atom.charge * (atom.origin[k] - self.center_of_charge[k]) for atom in self.atoms)
            self.mu[k] = -2 * np.trace(np.dot(self.P, self.M[k])) + nuclear
        self.mu *= 2.541765       # atomic units -> Debye

    def adj(self, x):
        assert x.shape[0] == x.shape[1]
        return np.conjugate(x).T

    def comm(self, A, B):
        return np.dot(A, B) - np.dot(B, A)

    # ---- DIIS --------------------------------------------------------------------------------
    def updateDIIS(self, F, P):
        FPS = multi_dot([F, P, self.S])
        err = multi_dot([self.X, FPS - self.adj(FPS), self.X])      # orthonormal-basis error vector
        self.fockSet.append(self.F)
        self.errorSet.append(err)
        if len(self.fockSet) > DIIS_DEPTH:
            del self.fockSet[0]
            del self.errorSet[0]
        m = len(self.fockSet)
        B = np.zeros((m + 1, m + 1))
        B[-1, :] = B[:, -1] = -1.0
        B[-1, -1] = 0.0
        for i in range(m):
            for j in range(i + 1):
                B[i, j] = B[j, i] = np.real(np.trace(np.dot(self.adj(self.errorSet[i]), self.errorSet[j])))
        rhs = np.zeros(m + 1)
        rhs[-1] = -1.0
        weights = np.linalg.solve(B, rhs)
        assert np.isclose(sum(weights[:-1]), 1.0)
        F_new = np.zeros((self.nbasis, self.nbasis), dtype="complex")
        for w, Fk in zip(weights, self.fockSet):
            F_new += w * Fk
        return F_new
