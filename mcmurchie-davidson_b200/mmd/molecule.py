"""Molecule: geometry + basis -> basis functions -> integrals.  Public behaviour follows the
reference's mmd/molecule.py (same constructor, attributes and method names); the integral work is
done by the B200 engine:

    build(direct=False)  ->  one-electron matrices (device kernel), then either the dense TwoE
                             tensor (doERIs) or the Schwarz table for the direct Fock build.
"""
import itertools
import os
import sys

import numpy as np
from scipy.linalg import fractional_matrix_power

from mmd._b200 import basisio
from mmd._b200.engine import owned_engine as engine_for
from mmd.integrals.twoe import Basis, doERIs
from mmd.scf import SCF

BOHR_PER_ANGSTROM = 1.0 / 0.52917721092      # mmd/molecule.py:224
AMU_TO_AU = 1822.8885

# isotope-averaged masses as tabulated by the reference (index = atomic number; the value for
# oxygen is the reference's, kept for drop-in behaviour)
_MASSES = [0.0, 1.008, 4.003, 6.941, 9.012, 10.812, 12.011, 14.007, 5.999, 18.998, 20.180, 22.990,
           24.305, 26.982, 28.086, 30.974, 32.066, 35.453, 39.948]

_SHELL_COMPONENTS = {
    "S": [(0, 0, 0)],
    "P": [(1, 0, 0), (0, 1, 0), (0, 0, 1)],
    "D": [(2, 0, 0), (1, 1, 0), (1, 0, 1), (0, 2, 0), (0, 1, 1), (0, 0, 2)],
    "F": [(3, 0, 0), (2, 1, 0), (2, 0, 1), (1, 2, 0), (1, 1, 1), (1, 0, 2), (0, 3, 0), (0, 2, 1), (0, 1, 2), (0, 0, 3)],
}


class Atom(object):
    def __init__(self, charge, mass, origin=np.zeros(3)):
        self.charge = charge
        self.mass = mass
        self.origin = origin
        self.forces = np.zeros(3)
        self.saved_forces = np.zeros(3)
        self.velocities = np.zeros(3)


class Molecule(SCF):
    """Molecule(geometry, basis='sto-3g'); geometry = 'charge mult' line then 'Sym x y z' (Angstrom)."""

    def __init__(self, geometry, basis="sto-3g"):
        self.charge, self.multiplicity, self.atoms = self.read_molecule(geometry)
        self.nelec = sum(atom.charge for atom in self.atoms) - self.charge
        self.nocc = self.nelec // 2
        self.is_built = False
        self.geometry_input = geometry
        self.basis_name = str(basis)
        self.basis_data = self.getBasis(basis)
        self.formBasis()

    # ---- parsing -----------------------------------------------------------------------------
    def sym2num(self, sym):
        return basisio.atomic_number(sym)

    def getBasis(self, filename):
        """{atomic number: [(momentum, [(exp, coef), ...]), ...]}; accepts a basis name or a path."""
        return basisio.load_basis(filename, os.path.join(os.path.dirname(os.path.abspath(__file__)), "basis"))

    def momentum2shell(self, momentum):
        return _SHELL_COMPONENTS[str(momentum)]

    def read_molecule(self, geometry):
        lines = [ln for ln in geometry.split("\n") if ln]
        head = lines[0].split()
        assert len(head) == 2
        charge, multiplicity = int(head[0]), int(head[1])
        atoms = []
        for ln in lines[1:]:
            tok = ln.split()
            if len(tok) == 0:
                break
            assert len(tok) == 4
            z = self.sym2num(tok[0])
            xyz = np.asarray([float(tok[1]) / 0.52917721092, float(tok[2]) / 0.52917721092,
                              float(tok[3]) / 0.52917721092])
            atoms.append(Atom(charge=z, mass=_MASSES[z] * AMU_TO_AU, origin=xyz))
        return charge, multiplicity, atoms

    # ---- basis functions ---------------------------------------------------------------------
    def formBasis(self):
        """self.bfs: atoms in input order -> shells in file order -> Cartesian components."""
        self.bfs = []
        self._engine = None                # the engine is tied to these Basis objects
        for atom in self.atoms:
            atom_first = len(self.bfs)
            for momentum, prims in self.basis_data[atom.charge]:
                if str(momentum) not in ("S", "P", "D"):
                    # fail here, with the reason, rather than deep inside the engine (the reference's general-L recursion
                    # accepts F shells; the device kernels are instantiated for (ss|ss) ... (dd|dd))
                    raise NotImplementedError("basis '%s' has %s functions on Z=%d: the B200 engine covers s, p and d shells "
                                              "((ss|ss) ... (dd|dd) class kernels; f-type Hermite tables exist only inside the "
                                              "gradient kernels)" % (self.basis_name, momentum, atom.charge))
                exps = np.asarray([e for e, _ in prims])
                coefs = np.asarray([c for _, c in prims])
                for lmn in self.momentum2shell(momentum):
                    self.bfs.append(Basis(np.asarray(atom.origin), np.asarray(lmn), len(exps), exps, coefs))
            atom._bf_range = (atom_first, len(self.bfs))
        self.nbasis = len(self.bfs)
        for atom in self.atoms:            # masks used by geometric-derivative code
            atom.mask = np.zeros(self.nbasis)
            atom.mask[atom._bf_range[0]:atom._bf_range[1]] = 1.0
        zsum = float(sum(atom.charge for atom in self.atoms))
        self.center_of_charge = np.asarray(
            [sum(atom.charge * atom.origin[k] for atom in self.atoms) for k in range(3)]) * (1.0 / zsum)

    # ---- integrals ---------------------------------------------------------------------------
    @property
    def engine(self):
        """This molecule's own device engine (shell-pair tables, Schwarz data, the resident TwoE): created on first
        use, dropped by formBasis when the basis functions change.  Owned here, so no other molecule and no
        element-wise ERI/S/T/V call can evict it."""
        eng = getattr(self, "_engine", None)
        key = tuple(id(b) for b in self.bfs)
        if eng is None or getattr(self, "_engine_key", None) != key:
            eng = engine_for(self.bfs)
            self._engine, self._engine_key = eng, key
        return eng

    def build(self, direct=False):
        self.one_electron_integrals()
        if direct:
            self.screen = self.engine.schwarz()     # dict keyed p(p+1)//2+q, like the reference
        else:
            self.two_electron_integrals()
        self.is_built = True

    def one_electron_integrals(self):
        Z = np.asarray([atom.charge for atom in self.atoms], dtype=float)
        xyz = np.asarray([atom.origin for atom in self.atoms], dtype=float)
        self.S, self.T, self.V, self.M, self.L = self.engine.onee(Z, xyz, self.center_of_charge)
        self.mu = np.zeros(3, dtype="complex")
        self.nuc_energy = 0.0
        for a, b in itertools.combinations(self.atoms, 2):
            self.nuc_energy += a.charge * b.charge / np.linalg.norm(a.origin - b.origin)
        self.Core = self.T + self.V
        self.X = fractional_matrix_power(self.S, -0.5)
        self.U = fractional_matrix_power(self.S, 0.5)

    def two_electron_integrals(self):
        """Dense (N,N,N,N) tensor (cython/twoe.pyx:12-31).  The B200 engine fills it ON THE DEVICE, where the in-core
        J/K, the AO->MO transformation and MP2 read it; the host copy behind `mol.TwoE` (1.66 GB for benzene/6-31G**,
        seconds of page-able copies) is only made when somebody reads that attribute."""
        fill = getattr(self.engine, "dense_device", None)
        if fill is not None and fill():
            self._TwoE = None
            self._TwoE_on_device = True
        else:
            N = self.nbasis
            self.TwoE = np.asarray(doERIs(N, np.zeros((N, N, N, N)), self.bfs))

    @property
    def TwoE(self):
        if getattr(self, "_TwoE", None) is None:
            if not getattr(self, "_TwoE_on_device", False):
                raise AttributeError("TwoE")          # like the reference before build(direct=False): hasattr() is False
            self._TwoE = np.ascontiguousarray(self.engine.dense_host())
        return self._TwoE

    @TwoE.setter
    def TwoE(self, value):
        self._TwoE = value
        self._TwoE_on_device = False

    @TwoE.deleter
    def TwoE(self):
        self._TwoE = None
        self._TwoE_on_device = False

    def forces(self):
        """Nuclear forces of the converged RHF state (mmd/forces.py of the reference): atom.forces = -dE/dX for every
        atom.  Derivative one- and two-electron integrals are evaluated on the device (csrc/grad.cu) and contracted
        with the densities as they are produced — the reference's N^4 derivative tensor per atom and direction is
        never formed, so forces also work after a direct SCF (no mol.TwoE needed)."""
        if not getattr(self, "is_converged", False):
            sys.exit("Need to converge SCF before computing gradient")
        atom_of_function = np.zeros(self.nbasis, dtype=np.int64)
        for k, atom in enumerate(self.atoms):
            atom_of_function[atom._bf_range[0]:atom._bf_range[1]] = k
        Z = [atom.charge for atom in self.atoms]
        xyz = [atom.origin for atom in self.atoms]
        g1, g2, gn = self.engine.gradient(Z, xyz, atom_of_function, self.P, self.F)
        self.gradient_parts = {"one_electron": g1, "two_electron": g2, "nuclear": gn}
        total = g1 + g2 + gn
        for k, atom in enumerate(self.atoms):
            atom.forces = -total[k]              # strictly dE/dX was computed; F = -dE/dX
        return self._forces

    @property
    def _forces(self):
        """(natom, 3) array of the forces last computed (mmd/molecule.py:48-53 of the reference)."""
        return np.concatenate([atom.forces for atom in self.atoms]).reshape(-1, 3)

    def save_integrals(self, folder=None):
        """Crawford-format text dump (enuc, nbf, nelec, s, t, v, eri with 1-based indices)."""
        if folder is None:
            sys.exit("Please provide a folder to save the integrals.")
        if not self.is_built:
            self.build()
        os.makedirs(folder, exist_ok=True)
        np.savetxt(folder + "/enuc.dat", np.asarray(self.nuc_energy).reshape(1,))
        np.savetxt(folder + "/nbf.dat", np.asarray(self.nbasis, dtype=int).reshape(1,), fmt="%d")
        np.savetxt(folder + "/nelec.dat", np.asarray(self.nelec, dtype=int).reshape(1,), fmt="%d")
        np.savetxt(folder + "/s.dat", self.S)
        np.savetxt(folder + "/t.dat", self.T)
        np.savetxt(folder + "/v.dat", self.V)
        n = self.nbasis
        with open(folder + "/eri.dat", "w") as f:
            for i, j, k, l in itertools.product(range(n), repeat=4):
                print(i + 1, j + 1, k + 1, l + 1, self.TwoE[i, j, k, l], file=f)
        with open(folder + "/geometry.txt", "w") as f:
            print(self.geometry_input, file=f)
