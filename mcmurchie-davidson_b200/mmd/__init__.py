"""mmd — drop-in Python surface of jjgoings/McMurchie-Davidson with the two-electron hot path
(ERIs + closed-shell J/K Fock build) running on NVIDIA B200 through libmmdb200.so.

Same entry points as the reference: Molecule(geometry, basis).RHF(), mol.TwoE,
mmd.integrals.twoe.{Basis, ERI, doERIs}, mmd.integrals.fock.formPT, PostSCF(mol).MP2().
"""
