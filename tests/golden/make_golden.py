#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Runs only in the build container: needs oracle/_ref (the reference built by oracle/build_ref.py)
and imports the reference's own `mmd` package from there — never the drop-in package.  The
fixtures it writes are what travels to the GPU box.

    PYTHONPATH=oracle/_ref python tests/golden/make_golden.py [--fast]

Fixtures
    anchors.json            SCF / MP2 energies, iteration counts and (E, RMS(P)) trajectories
    h2o_sto3g.npz           full TwoE, S T V M L, Schwarz table, formPT matrices
    h2o_ccpvdz.npz          unique TwoE (i>=j,k>=l,ij>=kl order), S T V M L, Schwarz, formPT matrices
    classes81.npz           the 81 s/p/d ERI class combinations of the reference's tests/test013.py
    sampled_<config>.npz    random basis-function quartets of the benchmark configurations with the
                            reference's ERI values (stratified over angular-momentum classes)
    boys_kat.json           the Boys known-answer table of the reference's tests/test012.py
    grad_h2o.npz            nuclear-derivative integrals of the reference's cython/grad.pyx (ERIx, Sx, Tx, VxA,
                            VxB) on random function tuples of H2O/STO-3G and H2O/cc-pVDZ — pins the oracle's
                            restatement of the NEXT scope row; `--grad-only` writes just this
"""
import importlib.util
import io
import json
import os
import re
import sys
import time
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from mmd.molecule import Molecule            # noqa: E402  (the REFERENCE)
from mmd.postscf import PostSCF              # noqa: E402
from mmd.integrals.twoe import ERI, Basis    # noqa: E402
from mmd.integrals.fock import formPT        # noqa: E402

assert os.path.realpath(sys.modules["mmd.molecule"].__file__).startswith(os.path.realpath(REF))

spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "mcmurchie-davidson_b200", "mmd", "_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

from oracle import oracle as O               # noqa: E402  (C restatement, used only to pick samples)

FAST = "--fast" in sys.argv


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with redirect_stdout(buf):
        r = fn(*a, **k)
    return r, buf.getvalue()


def run_scf(geom, basis, direct, conver=1e-8, mp2=False):
    mol = Molecule(geometry=geom, basis=basis)
    hist = []
    orig = mol.computeEnergy

    def hook():
        orig()
        hist.append(float(np.real(mol.energy)))
    mol.computeEnergy = hook
    _, out = quiet(mol.RHF, direct=direct, conver=conver)
    m = re.search(r"in (\d+) iterations", out)
    rec = {"energy": float(np.real(mol.energy)), "iterations": int(m.group(1)) if m else None,
           "converged": bool(mol.is_converged), "energies": hist, "P_RMS_final": float(np.real(mol.P_RMS)),
           "nbasis": int(mol.nbasis), "dipole": [float(np.real(x)) for x in mol.mu]}
    if mp2:
        quiet(PostSCF(mol).MP2)
        rec["emp2"] = float(np.real(mol.emp2))
    return mol, rec


def core_guess_density(mol):
    import scipy.linalg
    FO = mol.X.T @ mol.Core @ mol.X
    e, CO = scipy.linalg.eigh(FO)
    C = mol.X @ CO
    occ = C[:, :mol.nocc]
    return (occ @ occ.conj().T).astype("complex")


def unique_pack(T):
    N = T.shape[0]
    vals = []
    for i in range(N):
        for j in range(i + 1):
            ij = i * (i + 1) // 2 + j
            for k in range(N):
                for l in range(k + 1):
                    if ij >= k * (k + 1) // 2 + l:
                        vals.append(T[i, j, k, l])
    return np.asarray(vals)


def small_molecule_fixture(name, geom, basis, pack):
    t0 = time.time()
    mol, rec = run_scf(geom, basis, direct=False, mp2=True)
    N = mol.nbasis
    screen = {}
    for p in range(N):
        for q in range(p + 1):
            screen[p * (p + 1) // 2 + q] = ERI(mol.bfs[p], mol.bfs[q], mol.bfs[p], mol.bfs[q])
    scr = np.asarray([screen[k] for k in range(N * (N + 1) // 2)])
    P1 = core_guess_density(mol)
    Z = np.zeros_like(P1)
    G1 = formPT(P1, Z, mol.bfs, N, screen, 1e-12)
    # late-iteration-like incremental build: converged density vs a slightly perturbed one
    rng = np.random.default_rng(7)
    Pc = mol.P.astype("complex")
    pert = rng.standard_normal((N, N)) * 1e-7
    Pold = Pc - (pert + pert.T)
    G2 = formPT(Pc, Pold, mol.bfs, N, screen, 1e-12)
    # complex hermitian density (real-time TDHF style consumer)
    A = rng.standard_normal((N, N)) * 1e-2
    Pz = Pc + 1j * (A - A.T)
    G3 = formPT(Pz, Z, mol.bfs, N, screen, 1e-12)
    J = np.einsum("pqrs,sr->pq", mol.TwoE.astype("complex"), Pz)
    K = np.einsum("psqr,sr->pq", mol.TwoE.astype("complex"), Pz)
    np.savez_compressed(os.path.join(HERE, name + ".npz"),
                        TwoE=(unique_pack(mol.TwoE) if pack else mol.TwoE), packed=pack, N=N,
                        S=mol.S, T=mol.T, V=mol.V, M=mol.M, L=mol.L, screen=scr,
                        P1=P1, G1=G1, Pc=Pc, Pold=Pold, G2=G2, Pz=Pz, G3=G3, J3=J, K3=K,
                        C=mol.C, MO=mol.MO, energy=rec["energy"], emp2=rec["emp2"], nuc=mol.nuc_energy)
    print(name, "done in %.1fs" % (time.time() - t0))
    return rec


def classes81():
    Se = [3047.5249000, 457.3695100, 103.9486900, 29.2101550, 9.2866630, 3.1639270]
    Sc = [0.0018347, 0.0140373, 0.0688426, 0.2321844, 0.4679413, 0.3623120]
    Pe = [7.8682724, 1.8812885, 0.5442493]
    Pc = [0.0689991, 0.3164240, 0.7443083]
    De, Dc = [0.8], [1.0]
    s1 = Basis([0, 0, 0], (0, 0, 0), 6, Se, Sc); p1x = Basis([0, 0, 0], (1, 0, 0), 3, Pe, Pc); d1xx = Basis([0, 0, 0], (2, 0, 0), 1, De, Dc)
    s2 = Basis([2, 0, 0], (0, 0, 0), 6, Se, Sc); p2x = Basis([2, 0, 0], (1, 0, 0), 3, Pe, Pc); d2xx = Basis([2, 0, 0], (2, 0, 0), 1, De, Dc)
    vals = []
    for a in [s1, p2x, d1xx]:
        for b in [s2, p2x, d1xx]:
            for c in [s2, p1x, d2xx]:
                for d in [s1, p1x, d2xx]:
                    vals.append(ERI(a, b, c, d))
    # plus a less symmetric set: off-axis centres, mixed components
    rng = np.random.default_rng(3)
    fns, desc = [], []
    comps = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (2, 0, 0), (1, 1, 0), (1, 0, 1), (0, 2, 0), (0, 1, 1), (0, 0, 2)]
    cents = [[0.0, 0.0, 0.0], [1.1, -0.4, 0.7], [-0.6, 0.9, 1.3]]
    for ce in cents:
        for lmn in comps:
            L = sum(lmn)
            e, c = [(Se, Sc), (Pe, Pc), (De, Dc)][L]
            fns.append(Basis(ce, lmn, len(e), e, c))
            desc.append((ce, lmn))
    idx = rng.integers(0, len(fns), size=(400, 4))
    v2 = [ERI(fns[i], fns[j], fns[k], fns[l]) for i, j, k, l in idx]
    np.savez_compressed(os.path.join(HERE, "classes81.npz"), vals81=np.asarray(vals), idx=idx, vals=np.asarray(v2),
                        centres=np.asarray(cents), comps=np.asarray(comps))
    print("classes81 done")


def sampled(name, nper):
    geom, basis = synth.config(name)
    mol = Molecule(geometry=geom, basis=basis)
    N = mol.nbasis
    fb = O.FlatBasis(mol.bfs)
    Q = O.schwarz(fb)               # oracle C (validated against the reference) — only to choose samples
    sq = np.sqrt(np.abs(Q))
    L = np.asarray([int(sum(b.shell)) for b in mol.bfs])
    rng = np.random.default_rng(0)
    p, q = np.tril_indices(N)
    picks = []
    # strata by (La,Lb,Lc,Ld); draw significant pairs so the integrals are not all ~0
    for la in range(3):
        for lb in range(la + 1):
            pa = np.nonzero((L[p] == la) & (L[q] == lb) | (L[p] == lb) & (L[q] == la))[0]
            pa = pa[sq[pa] > 1e-4]
            for lc in range(3):
                for ld in range(lc + 1):
                    pb = np.nonzero((L[p] == lc) & (L[q] == ld) | (L[p] == ld) & (L[q] == lc))[0]
                    pb = pb[sq[pb] > 1e-4]
                    if len(pa) == 0 or len(pb) == 0:
                        continue
                    ia = rng.choice(pa, size=nper)
                    ib = rng.choice(pb, size=nper)
                    for x, y in zip(ia, ib):
                        picks.append((p[x], q[x], p[y], q[y]))
    idx = np.asarray(picks, dtype=np.int64)
    t0 = time.time()
    vals = np.asarray([ERI(mol.bfs[i], mol.bfs[j], mol.bfs[k], mol.bfs[l]) for i, j, k, l in idx])
    dt = time.time() - t0
    # Schwarz diagonal samples from the reference as well
    sidx = rng.integers(0, len(p), size=200)
    svals = np.asarray([ERI(mol.bfs[p[s]], mol.bfs[q[s]], mol.bfs[p[s]], mol.bfs[q[s]]) for s in sidx])
    np.savez_compressed(os.path.join(HERE, "sampled_%s.npz" % name), idx=idx, vals=vals, N=N,
                        schwarz_pq=np.stack([p[sidx], q[sidx]], 1), schwarz_vals=svals, ref_seconds=dt)
    print("sampled", name, len(idx), "quartets, reference ERI time %.1fs" % dt)


def boys_kat():
    src = open(os.path.join(REF, "tests", "test012.py")).read()
    kats = re.findall(r"boys\(\s*([\d.Ee+-]+)\s*,\s*([\d.Ee+-]+)\s*\)\s*,\s*([\d.Ee+-]+)", src)
    json.dump([[float(a), float(b), float(c)] for a, b, c in kats], open(os.path.join(HERE, "boys_kat.json"), "w"))
    print("boys KATs:", len(kats))


def grad_fixture():
    """Derivative integrals straight from the reference's compiled grad module."""
    from mmd.integrals import grad as G
    out = {}
    for tag, basis in (("sto3g", "sto-3g"), ("ccpvdz", "cc-pvdz")):
        mol = Molecule(geometry=synth.water(), basis=basis)
        bfs, N = mol.bfs, mol.nbasis
        rng = np.random.default_rng(7)
        n4, n2 = (200, 150) if tag == "sto3g" else (300, 200)
        idx = rng.integers(0, N, size=(n4, 4))
        xs, cs = rng.integers(0, 3, size=n4), rng.integers(0, 4, size=n4)
        out[tag + "_eri_idx"], out[tag + "_eri_x"], out[tag + "_eri_c"] = idx, xs, cs
        out[tag + "_eri"] = np.array([G.ERIx(bfs[i], bfs[j], bfs[k], bfs[l], x=int(x), center="abcd"[c])
                                      for (i, j, k, l), x, c in zip(idx, xs, cs)])
        pr = rng.integers(0, N, size=(n2, 2))
        x2, c2, at = rng.integers(0, 3, size=n2), rng.integers(0, 2, size=n2), rng.integers(0, len(mol.atoms), size=n2)
        out[tag + "_pr"], out[tag + "_x2"], out[tag + "_c2"], out[tag + "_atom"] = pr, x2, c2, at
        out[tag + "_atoms_xyz"] = np.array([a.origin for a in mol.atoms])
        out[tag + "_S"] = np.array([G.Sx(bfs[i], bfs[j], x=int(x), center="AB"[c]) for (i, j), x, c in zip(pr, x2, c2)])
        out[tag + "_T"] = np.array([G.Tx(bfs[i], bfs[j], x=int(x), center="AB"[c]) for (i, j), x, c in zip(pr, x2, c2)])
        out[tag + "_VA"] = np.array([G.VxA(bfs[i], bfs[j], np.asarray(mol.atoms[a].origin), x=int(x))
                                     for (i, j), x, a in zip(pr, x2, at)])
        out[tag + "_VB"] = np.array([G.VxB(bfs[i], bfs[j], np.asarray(mol.atoms[a].origin), x=int(x), center="AB"[c])
                                     for (i, j), x, c, a in zip(pr, x2, c2, at)])
    # RHF forces of the reference (mmd/forces.py) with the converged P and F they were computed from
    h2 = "\n0 1\nH 0.0 0.0 0.74\nH 0.0 0.0 0.0\n"
    for tag, geom, basis in (("h2", h2, "sto-3g"), ("h2o", synth.water(), "sto-3g"), ("h2o_ccpvdz", synth.water(), "cc-pvdz")):
        mol = Molecule(geometry=geom, basis=basis)      # cc-pVDZ: d functions -> f-type shifted integrals (minutes in the reference)
        quiet(mol.RHF)
        quiet(mol.forces)
        out["forces_" + tag] = np.array([a.forces for a in mol.atoms])
        out["forces_" + tag + "_P"], out["forces_" + tag + "_F"] = np.asarray(mol.P), np.asarray(mol.F)
    np.savez_compressed(os.path.join(HERE, "grad_h2o.npz"), **out)
    print("grad fixture:", {k: v.shape for k, v in out.items() if k.endswith(("_eri", "_S"))})


def main():
    if "--grad-only" in sys.argv:
        grad_fixture()
        return
    boys_kat()
    grad_fixture()
    anchors = {}
    anchors["h2o_sto3g_incore"] = small_molecule_fixture("h2o_sto3g", synth.water(), "sto-3g", pack=False)
    anchors["h2o_ccpvdz_incore"] = small_molecule_fixture("h2o_ccpvdz", synth.water(), "cc-pvdz", pack=True)
    classes81()
    h2 = "\n0 1\nH 0.0 0.0 0.74\nH 0.0 0.0 0.0\n"
    he2 = "\n0 1\nHe 0.0 0.0 0.0\nHe 0.0 0.0 3.0\n"
    runs = [("h2_sto3g_incore", h2, "sto-3g", False, 1e-8, False),
            ("h2o_sto3g_direct", synth.water(), "sto-3g", True, 1e-8, False),
            ("h2o_sto3g_incore_tight", synth.water(), "sto-3g", False, 1e-14, False),
            ("ch4_sto3g_incore", synth.methane(), "sto-3g", False, 1e-8, False),
            ("ch4_sto3g_direct", synth.methane(), "sto-3g", True, 1e-8, False),
            ("ch4_321g_incore", synth.methane(), "3-21g", False, 1e-8, True),
            ("he2_ccpvdz_incore", he2, "cc-pvdz", False, 1e-8, True),
            ("h2o_dz_incore", synth.water(), "dz", False, 1e-8, False),
            ("h2o_321g_incore", synth.water(), "3-21g", False, 1e-8, False),
            ("h2o_631ppgss_incore", synth.water(), "6-31ppgss", False, 1e-8, False)]
    if not FAST:
        runs.append(("h2o_ccpvdz_direct", synth.water(), "cc-pvdz", True, 1e-8, False))
    for name, g, b, direct, conv, mp2 in runs:
        t0 = time.time()
        _, rec = run_scf(g, b, direct, conv, mp2)
        rec["geometry"], rec["basis"], rec["direct"], rec["conver"] = g, b, direct, conv
        anchors[name] = rec
        print(name, rec["energy"], rec["iterations"], "%.1fs" % (time.time() - t0))
    json.dump(anchors, open(os.path.join(HERE, "anchors.json"), "w"), indent=1)
    for name, nper in (("benzene_631gss", 12), ("w8_ccpvdz", 12), ("c20h42_631gs", 8), ("w32_ccpvdz", 8)):
        sampled(name, 3 if FAST else nper)


if __name__ == "__main__":
    main()
