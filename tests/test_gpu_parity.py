"""Parity of the CUDA path (through the ctypes C-ABI binding) against the CPU oracle, the golden
vectors produced by the reference, and — at the benchmark sizes the oracle cannot reach — through
size-independent properties (direct build == in-core build, shard additivity, linearity, 8-fold
symmetry).  Tolerances are the north star's: ERIs 1e-12 abs, Fock 1e-10, energies 1e-9 Eh."""
import os
import numpy as np
import pytest

from conftest import unique_pack
from mmd._b200 import engine as E
from mmd._b200 import synth
from mmd.integrals.twoe import ERI, Basis, doERIs
from mmd.molecule import Molecule
from test_oracle_pinned import _class_functions

pytestmark = pytest.mark.gpu

ERI_TOL = 1e-12
FOCK_TOL = 1e-10
E_TOL = 1e-9


def test_boys_device_vs_oracle(oracle):
    rng = np.random.default_rng(0)
    Ts = np.concatenate([10 ** rng.uniform(-9, 6, 3000), rng.uniform(0, 45, 3000), [0.0, 39.999999, 40.0, 40.000001, 59.9999, 60.0, 60.0001, 119.9, 120.1]])
    for n in (0, 2, 5, 8):
        got = E.boys(n, Ts)
        ref = np.array([[oracle.boys(m, T) for m in range(n + 1)] for T in Ts])
        assert (np.abs(got - ref) / np.abs(ref)).max() < 5e-15


def test_templated_boys_of_the_class_kernels_vs_oracle(oracle):
    """The routine the class kernels actually run (prim_Fs<L>, csrc/core.cuh): per L its own Taylor-table range and
    its own switch-over T_max(L) to the alpha-free asymptotic series — dense samples on both sides of every switch."""
    tmax = [37, 41, 44, 47, 50, 53, 55, 58, 60]
    rng = np.random.default_rng(1)
    for Ltot in range(9):
        tm = tmax[Ltot]
        Ts = np.concatenate([10 ** rng.uniform(-9, 6, 2000), rng.uniform(0, tm, 3000), rng.uniform(tm - 2, tm + 2, 2000),
                             rng.uniform(tm, 130, 1000), np.arange(0, 8 * tm + 1) / 8.0, np.arange(0, 8 * tm) / 8.0 + 0.0625,
                             [0.0, tm - 1e-9, float(tm), tm + 1e-9, np.nextafter(tm, 0), np.nextafter(tm, 1e9)]])
        got = E.boys_class(Ltot, Ts)
        ref = np.array([[oracle.boys(m, T) for m in range(Ltot + 1)] for T in Ts])
        rel = np.abs(got - ref) / np.abs(ref)
        assert rel.max() < 5e-15, (Ltot, float(Ts[rel.max(axis=1).argmax()]), float(rel.max()))


@pytest.mark.parametrize("cfg", ["h2o_sto3g", "h2o_ccpvdz"])
def test_dense_tensor(oracle, golden, cfg):
    g = golden(cfg + ".npz")
    mol = Molecule(*synth.config(cfg))
    N = mol.nbasis
    T = np.zeros((N,) * 4)
    out = doERIs(N, T, mol.bfs)
    assert out is T
    ref = np.zeros((N,) * 4)
    oracle.doERIs(N, ref, mol.bfs)
    assert np.abs(T - ref).max() < ERI_TOL
    gold = g["TwoE"]
    assert np.abs((unique_pack(T) if bool(g["packed"]) else T) - gold).max() < ERI_TOL


def test_81_class_combinations_and_handbuilt_functions(golden):
    g = golden("classes81.npz")
    fns, combos = _class_functions(g)
    got = np.array([ERI(*q) for q in combos])
    assert np.abs(got - g["vals81"]).max() < ERI_TOL
    eng = E.engine_for(fns)
    assert np.abs(eng.eri_quartets(g["idx"]) - g["vals"]).max() < ERI_TOL
    assert np.abs(eng.eri_quartets(g["idx"], impl=1) - g["vals"]).max() < ERI_TOL      # generic kernel
    with pytest.raises(TypeError):
        ERI(fns[0], fns[1], fns[2], "not a basis function")


def test_partial_shell_lists_dense(oracle):
    # a list that is NOT made of complete shells: ghost components must not leak into the tensor
    mk = lambda lmn, c, e: Basis(c, lmn, len(e), e, [1.0] * len(e))
    bfs = [mk((1, 1, 0), [0, 0, 0], [0.8]), mk((0, 0, 0), [0, 0, 0], [1.3, 0.4]), mk((0, 0, 1), [0.5, 0.2, 1.0], [0.9]),
           mk((2, 0, 0), [0.5, 0.2, 1.0], [0.6]), mk((0, 1, 0), [-0.7, 0.1, 0.3], [1.1, 0.3])]
    N = len(bfs)
    T = np.zeros((N,) * 4)
    doERIs(N, T, bfs)
    ref = np.zeros((N,) * 4)
    oracle.doERIs(N, ref, bfs)
    assert np.abs(T - ref).max() < ERI_TOL


@pytest.mark.parametrize("cfg", ["benzene_631gss", "w8_ccpvdz", "c20h42_631gs", "w32_ccpvdz"])
def test_sampled_quartets_benchmark_configs(golden, cfg):
    g = golden("sampled_%s.npz" % cfg)
    mol = Molecule(*synth.config(cfg))
    eng = mol.engine
    assert np.abs(eng.eri_quartets(g["idx"]) - g["vals"]).max() < ERI_TOL
    assert np.abs(eng.eri_quartets(g["idx"], impl=1) - g["vals"]).max() < ERI_TOL
    scr = eng.schwarz()
    pq = g["schwarz_pq"]
    got = scr.flat[pq[:, 0] * (pq[:, 0] + 1) // 2 + pq[:, 1]]
    assert np.abs(got - g["schwarz_vals"]).max() < ERI_TOL


@pytest.mark.parametrize("cfg", ["benzene_631gss", "w8_ccpvdz", "c20h42_631gs", "w32_ccpvdz"])
def test_big_stratified_samples_benchmark_configs(golden, cfg):
    """The reference's own ERI values on >= 1e4 function quartets per class (2.5e3 for configurations 2 and 3), five
    pair-distance bins per class: elementwise <= 1e-12 for the class kernels AND the generic kernel."""
    g = golden("sampled_big_%s.npz" % cfg)
    mol = Molecule(*synth.config(cfg))
    eng = mol.engine
    idx = g["idx"].astype(np.int64)
    err = np.abs(eng.eri_quartets(idx) - g["vals"])
    assert err.max() < ERI_TOL, (int(err.argmax()), idx[err.argmax()], float(err.max()))
    sub = slice(0, None, 7)
    assert np.abs(eng.eri_quartets(idx[sub], impl=1) - g["vals"][sub]).max() < ERI_TOL


@pytest.mark.parametrize("cfg", ["h2o_sto3g", "h2o_ccpvdz"])
def test_schwarz_formPT_jk_onee(oracle, golden, cfg):
    g = golden(cfg + ".npz")
    mol = Molecule(*synth.config(cfg))
    N = mol.nbasis
    eng = mol.engine
    scr = eng.schwarz()
    assert isinstance(scr, dict) and len(scr) == N * (N + 1) // 2
    assert np.abs(scr.flat - g["screen"]).max() < ERI_TOL
    Z = np.zeros((N, N), dtype=complex)
    from mmd.integrals.fock import formPT
    for P, Po, key in ((g["P1"], Z, "G1"), (g["Pc"], g["Pold"], "G2"), (g["Pz"], Z, "G3")):
        G = formPT(P, Po, mol.bfs, N, scr, 1e-12)
        assert G.dtype == np.complex128 and G.shape == (N, N)
        assert np.abs(G - g[key]).max() < FOCK_TOL                       # un-symmetrised, elementwise
        assert np.abs(G - oracle.formPT(P, Po, mol.bfs, N, g["screen"], 1e-12)).max() < FOCK_TOL
    # caller-supplied plain dict (not the engine's own table) and a loose tolerance
    plain = dict((k, float(v) * 1.0) for k, v in enumerate(g["screen"]))
    G = formPT(g["P1"], Z, mol.bfs, N, plain, 1e-6)
    assert np.abs(G - oracle.formPT(g["P1"], Z, mol.bfs, N, g["screen"], 1e-6)).max() < FOCK_TOL
    eng.dense()
    J, K = eng.jk_incore(g["Pz"])
    assert np.abs(J - g["J3"]).max() < FOCK_TOL and np.abs(K - g["K3"]).max() < FOCK_TOL
    S, T, V, M, L = eng.onee([a.charge for a in mol.atoms], [a.origin for a in mol.atoms], mol.center_of_charge)
    for got, key in ((S, "S"), (T, "T"), (V, "V"), (M, "M"), (L, "L")):
        assert np.abs(got - g[key]).max() < 1e-12


def test_onee_larger_molecule_vs_oracle(oracle):
    # one-electron enabler at a size the golden fixtures do not cover (many atoms, large Boys arguments)
    mol = Molecule(*synth.config("benzene_631gss"))
    Z = [a.charge for a in mol.atoms]
    xyz = [a.origin for a in mol.atoms]
    got = mol.engine.onee(Z, xyz, mol.center_of_charge)
    ref = oracle.onee(mol.bfs, Z, xyz, mol.center_of_charge)
    for g, r in zip(got, ref):
        assert np.abs(g - r).max() < 1e-11


def test_host_buffer_c_abi_entry_points(oracle, golden):
    """The pure C-ABI host-buffer calls (what a non-Python binding would use, INTEGRATION.md option B):
    mmdb_schwarz_host, mmdb_set_schwarz_host, mmdb_eri_dense_host, mmdb_formPT_host — numpy in, numpy out."""
    import ctypes as C
    from mmd._b200 import lib as L
    g = golden("h2o_ccpvdz.npz")
    mol = Molecule(*synth.config("h2o_ccpvdz"))
    eng = mol.engine
    N = mol.nbasis
    lib = eng.lib
    Q = np.zeros(N * (N + 1) // 2)
    L.check(lib.mmdb_schwarz_host(eng.h, L.ptr(Q)))
    assert np.abs(Q - g["screen"]).max() < ERI_TOL
    T = np.zeros((N,) * 4)
    L.check(lib.mmdb_eri_dense_host(eng.h, L.ptr(T)))
    assert np.abs(unique_pack(T) - g["TwoE"]).max() < ERI_TOL
    scr_flat = np.ascontiguousarray(g["screen"], dtype=np.float64)       # keep alive across the call
    L.check(lib.mmdb_set_schwarz_host(eng.h, L.ptr(scr_flat)))
    for P, Po, key in ((g["P1"], np.zeros_like(g["P1"]), "G1"), (g["Pz"], np.zeros_like(g["Pz"]), "G3")):
        Pc = np.ascontiguousarray(P, dtype=np.complex128)
        Poc = np.ascontiguousarray(Po, dtype=np.complex128)
        G = np.zeros((N, N), dtype=np.complex128)
        stats = L.FockStats()
        L.check(lib.mmdb_formPT_host(eng.h, L.ptr(Pc.view(np.float64)), L.ptr(Poc.view(np.float64)), 1e-12,
                                     L.ptr(G.view(np.float64)), C.byref(stats)))
        assert np.abs(G - g[key]).max() < FOCK_TOL
        assert stats.quartets > 0 and stats.prim_quartets >= stats.quartets
    # error convention: non-zero return + message, no exception from C
    rc = lib.mmdb_eri_shell_quartets(eng.h, 0, 1, 1, None, None, None, 0, None)       # pc_bra < pc_ket is invalid
    assert rc != 0 and b"pc_bra" in lib.mmdb_last_error()


def test_c_abi_communicator_and_allreduce():
    """mmdb_comm_unique_id / mmdb_comm_init / mmdb_allreduce_G / mmdb_comm_destroy (NCCL behind the C ABI) with the
    single-rank communicator: the reduction is the identity and the sharded call with nshards = 1 is the whole build.
    Several ranks are one PROCESS per GPU: that path (Engine.formPT under torchrun -> dist.c_abi_comm -> mmdb_allreduce_G)
    is what bench.py --gpus N times as `e2e`.  Two ranks in two THREADS of one process were tried and dropped: NCCL's
    communicator setup dead-locks against the other thread's device allocations on a two-device box."""
    import ctypes as C
    import torch
    from mmd._b200 import lib as L
    lib = L.load()
    uid = (C.c_ubyte * 128)()
    L.check(lib.mmdb_comm_unique_id(uid))
    mol = Molecule(*synth.config("h2o_ccpvdz"))
    N = mol.nbasis
    rng = np.random.default_rng(3)
    A = rng.standard_normal((N, N))
    P = np.ascontiguousarray(0.1 * (A + A.T))
    eng = mol.engine
    eng.schwarz()
    full = eng.formPT(P.astype(complex), np.zeros((N, N), dtype=complex), tol=1e-12).real
    comm = C.c_void_p()
    L.check(lib.mmdb_comm_init(eng.device, 1, 0, uid, C.byref(comm)))
    dP = torch.from_numpy(P).to(eng.tdev)
    G = torch.zeros((N, N), dtype=torch.float64, device=eng.tdev)
    st = C.c_void_p(torch.cuda.current_stream(eng.tdev).cuda_stream)
    L.check(lib.mmdb_fock_direct(eng.h, L.ptr(dP), None, 1e-12, L.ptr(G), None, 0, 1, 0, None, st))
    L.check(lib.mmdb_allreduce_G(comm, L.ptr(G), G.numel(), 0, st))
    torch.cuda.synchronize(eng.tdev)
    assert np.abs(G.cpu().numpy() - full).max() < FOCK_TOL
    L.check(lib.mmdb_comm_destroy(comm))


def test_generally_contracted_s_shells_share_their_primitives():
    """cc-pVDZ's first two s functions of every heavy atom are two contractions of ONE primitive set.  The direct build
    evaluates them as a two-component pseudo-shell (S2): every primitive quartet once, contraction weights applied with
    the Hermite coefficients.  Same G as the build over the reference's own shells (mmdb_fock_direct flags bit2), fewer
    primitive quartets evaluated, and the reference-unit statistics of the grouped build are a superset count."""
    mol = Molecule(*synth.config("w8_ccpvdz"))
    dens = closed_form_densities(mol.bfs)
    eng = mol.engine
    scr = eng.schwarz()
    for name in ("A", "B"):
        P = dens[name].astype(complex)
        G_plain = eng.formPT(P, np.zeros_like(P), screen=scr, tol=1e-12, flags=4)
        st_plain = dict(eng.last_stats)
        G_gc = eng.formPT(P, np.zeros_like(P), screen=scr, tol=1e-12)
        st_gc = dict(eng.last_stats)
        assert np.abs(G_gc - G_plain).max() < FOCK_TOL
        assert st_plain["exec_prim_quartets"] == st_plain["prim_quartets"]
        assert st_gc["exec_prim_quartets"] < 0.8 * st_plain["prim_quartets"]
        assert st_gc["quartets"] >= st_plain["quartets"] and st_gc["prim_quartets"] >= st_plain["prim_quartets"]
    # tol = 0: nothing is screened
    P = dens["A"].astype(complex)
    G0p = eng.formPT(P, np.zeros_like(P), screen=scr, tol=0.0, flags=4)
    G0g = eng.formPT(P, np.zeros_like(P), screen=scr, tol=0.0)
    assert np.abs(G0g - G0p).max() < FOCK_TOL
    # complex (Hermitian) density and the deterministic mode run the per-function digestion of the grouped classes
    rng = np.random.default_rng(5)
    N = mol.nbasis
    A = rng.standard_normal((N, N)) * 0.01
    Pc = dens["A"] + 1j * (A - A.T)
    Gc_plain = eng.formPT(Pc, np.zeros_like(Pc), screen=scr, tol=1e-12, flags=4)
    Gc_gc = eng.formPT(Pc, np.zeros_like(Pc), screen=scr, tol=1e-12)
    assert np.abs(Gc_gc - Gc_plain).max() < FOCK_TOL


def test_far_and_near_lists_partition_the_work(monkeypatch):
    """MMDB_FAR_MAXL=3: the screening kernel sorts the block-digestible entries (one per slice of <= 8 bra primitive pairs)
    into a far-field list (every primitive quartet on the asymptotic Boys branch, proved from bounding spheres; kernels
    without Boys table) and a near list.  Same G as the default single-list build, the entries partition exactly, and the
    far list is populated.  (Off by default: see far_enabled in csrc/lib.cu for the measurement.)"""
    monkeypatch.setenv("MMDB_NO_GC", "1")       # the far lists exist for the plain shell classes only
    mol = Molecule(*synth.config("w8_ccpvdz"))
    P = closed_form_densities(mol.bfs)["A"].astype(complex)
    eng = mol.engine
    scr = eng.schwarz()
    G0 = eng.formPT(P, np.zeros_like(P), screen=scr, tol=1e-12)
    st0 = dict(eng.last_stats)
    assert st0["far_entries"] == 0 and st0["near_entries"] > 0
    monkeypatch.setenv("MMDB_FAR_MAXL", "3")
    G1 = eng.formPT(P, np.zeros_like(P), screen=scr, tol=1e-12)
    st1 = dict(eng.last_stats)
    assert st1["far_entries"] > 0 and st1["near_entries"] > 0
    assert st1["far_entries"] + st1["near_entries"] == st0["near_entries"]
    assert st1["quartets"] == st0["quartets"] and st1["prim_quartets"] == st0["prim_quartets"]
    assert np.abs(G1 - G0).max() < FOCK_TOL


def test_jk_incore_even_and_odd_sizes(oracle):
    rng = np.random.default_rng(2)
    he2 = "\n0 1\nHe 0.0 0.0 0.0\nHe 0.0 0.0 3.0\n"
    for geom, basis in ((he2, "cc-pvdz"), (synth.methane(), "sto-3g"), ("\n0 1\nH 0 0 0\nH 0 0 0.74\n", "sto-3g")):
        mol = Molecule(geom, basis)
        N = mol.nbasis
        T = mol.engine.dense()
        A = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
        for P in (A + A.conj().T, (A + A.conj().T).real.astype(complex)):
            J, K = mol.engine.jk_incore(P)
            Jr, Kr = oracle.jk_incore(T, P)
            assert np.abs(J - Jr).max() < FOCK_TOL and np.abs(K - Kr).max() < FOCK_TOL


ANCHORS = ["h2_sto3g_incore", "h2o_sto3g_incore", "h2o_sto3g_direct", "ch4_321g_incore",
           "he2_ccpvdz_incore", "h2o_dz_incore", "h2o_321g_incore", "h2o_631ppgss_incore", "h2o_ccpvdz_incore",
           "h2o_ccpvdz_direct"]


@pytest.mark.parametrize("name", ANCHORS)
def test_rhf_mp2_end_to_end_vs_reference(golden, name):
    from mmd.postscf import PostSCF
    a = golden("anchors.json")[name]
    geom = a.get("geometry", synth.water())
    basis = a.get("basis", "sto-3g" if "sto3g" in name else "cc-pvdz")
    direct = a.get("direct", name.endswith("direct"))
    mol = Molecule(geom, basis)
    mol.RHF(doPrint=False, direct=direct, conver=a.get("conver", 1e-8))
    assert mol.is_converged
    assert mol.scf_iterations == a["iterations"]
    assert abs(mol.energy.real - a["energy"]) < E_TOL
    if direct:
        assert not hasattr(mol, "TwoE")
    else:
        assert isinstance(mol.TwoE, np.ndarray) and mol.TwoE.flags.c_contiguous
    if "emp2" in a:
        PostSCF(mol).MP2()
        assert abs(mol.emp2.real - a["emp2"]) < E_TOL


@pytest.mark.parametrize("direct", [False, True])
def test_device_scf_loop_equals_host_scf_loop(monkeypatch, direct):
    """the RHF loop with device-resident linear algebra (default) against the NumPy/SciPy
    loop (MMDB_HOST_SCF=1) on the same GPU Fock builds: same iteration count, energy, density, orbital energies."""
    geom, basis = synth.config("h2o_ccpvdz")
    dev = Molecule(geom, basis)
    dev.RHF(doPrint=False, direct=direct)
    monkeypatch.setenv("MMDB_HOST_SCF", "1")
    host = Molecule(geom, basis)
    host.RHF(doPrint=False, direct=direct)
    assert dev.is_converged and host.is_converged
    assert dev.scf_iterations == host.scf_iterations
    assert abs(dev.energy.real - host.energy.real) < E_TOL
    assert np.abs(dev.P - host.P).max() < 1e-8 and np.abs(dev.F - host.F).max() < 1e-8
    assert np.abs(dev.MO - host.MO).max() < 1e-8
    for attr in ("P", "F", "C", "MO", "FO", "CO", "G", "P_old"):
        assert isinstance(getattr(dev, attr), np.ndarray), attr
    assert np.abs(np.asarray(dev.mu) - np.asarray(host.mu)).max() < 1e-6
    assert len(dev.scf_history) == len(host.scf_history)
    assert max(abs(a[0] - b[0]) for a, b in zip(dev.scf_history, host.scf_history)) < 1e-8


def test_tight_convergence_is_noise_limited(golden):
    """conver=1e-14 asks RMS(P) to drop below the rounding noise of the Fock build itself: the iteration
    count then depends on summation order (the device reductions use FP64 atomics), so only convergence
    and the energy are asserted.  (The oracle-backed CPU driver needs 14 iterations where the reference needs 20.)"""
    a = golden("anchors.json")["h2o_sto3g_incore_tight"]
    mol = Molecule(synth.water(), "sto-3g")
    mol.RHF(doPrint=False, conver=1e-14)
    assert mol.is_converged and abs(mol.energy.real - a["energy"]) < E_TOL


def test_deterministic_mode_is_bitwise_reproducible(oracle, golden):
    """MMDB_DETERMINISTIC: fixed-point integer accumulation -> identical bits run to run and for any shard split,
    still within the Fock tolerance of the oracle."""
    import ctypes as C
    import torch
    from mmd._b200 import lib as L
    g = golden("h2o_ccpvdz.npz")
    mol = Molecule(*synth.config("h2o_ccpvdz"))
    eng = mol.engine
    scr = eng.schwarz()
    eng.deterministic = True
    try:
        Z = np.zeros_like(g["P1"])
        G1 = eng.formPT(g["P1"], Z, screen=scr, tol=1e-12)
        G2 = eng.formPT(g["P1"], Z, screen=scr, tol=1e-12)
        assert np.array_equal(G1, G2)
        assert np.abs(G1 - g["G1"]).max() < FOCK_TOL
        # three shards accumulated as integers == the unsharded build, bit for bit
        N = mol.nbasis
        re = torch.from_numpy(np.ascontiguousarray(g["P1"].real)).to(eng.tdev)
        acc = torch.zeros((N, N), dtype=torch.float64, device=eng.tdev)
        for s in range(3):
            L.check(eng.lib.mmdb_fock_direct(eng.h, L.ptr(re), None, 1e-12, L.ptr(acc), None, s, 3, 2, None, eng._stream()))
        L.check(eng.lib.mmdb_fixed_to_double(eng.device, L.ptr(acc), acc.numel(), eng._stream()))
        assert np.array_equal(acc.cpu().numpy(), G1.real)
    finally:
        eng.deterministic = False


def test_ao2mo_mp2_device_vs_host(golden):
    """Device AO->MO (cuBLAS DGEMM quarter transformations) + MP2 kernel against the reference's formulas on the host."""
    from mmd.postscf import PostSCF
    for cfg in ("h2o_sto3g", "h2o_ccpvdz"):
        mol = Molecule(*synth.config(cfg))
        mol.RHF(doPrint=False)
        post = PostSCF(mol)
        assert post._e2_device is not None            # the device path ran
        C = np.real(mol.C)
        ref = np.einsum("pqrs,pP,qQ,rR,sS->PQRS", mol.TwoE, C, C, C, C, optimize=True)
        assert np.abs(mol.single_bar - ref).max() < 1e-11
        post.MP2()
        assert abs(mol.emp2.real - float(golden(cfg + ".npz")["emp2"])) < E_TOL
        dev = mol.emp2.real
        post._e2_device = None                         # force the host loop of the reference
        post.MP2()
        assert abs(mol.emp2.real - dev) < 1e-11


def test_degenerate_guess_case_ch4_sto3g(golden, monkeypatch):
    """CH4/STO-3G — the reference's only direct-SCF test (tests/test010.py:17-18) and tests/test003.py.  The core
    guess splits a degenerate t2 set between occupied and virtual orbitals, so the first density depends on how
    the eigensolver rotates the degenerate vectors; the reference's own two modes take 10 and 11 iterations.  The
    SCF loop diagonalises small matrices with the reference's own LAPACK call, so the trajectory is the
    reference's up to the rounding noise of the Fock build: |delta iterations| <= 1 and 1e-9 Eh, for the default
    loop and for the parity mode (host loop + deterministic fixed-point accumulation)."""
    for name in ("ch4_sto3g_incore", "ch4_sto3g_direct"):
        a = golden("anchors.json")[name]
        mol = Molecule(a["geometry"], a["basis"])
        mol.RHF(doPrint=False, direct=a["direct"])
        _ch4_check(mol, a, name)
    monkeypatch.setenv("MMDB_HOST_SCF", "1")
    monkeypatch.setenv("MMDB_DETERMINISTIC", "1")
    for name in ("ch4_sto3g_incore", "ch4_sto3g_direct"):
        a = golden("anchors.json")[name]
        mol = Molecule(a["geometry"], a["basis"])
        mol.RHF(doPrint=False, direct=a["direct"])
        _ch4_check(mol, a, name)


def _ch4_check(mol, a, name):
    """|delta iterations| <= 1 and 5e-8 Eh at the default stop, and the CONVERGED energy to 1e-9 Eh.
    The SCF stops at RMS(P) < 1e-8, where the (lagging, non-variational) energy expression still moves by O(1e-8) per
    iteration, and this trajectory is noise-limited (a degenerate t2 set straddles the occupied/virtual boundary of the
    core guess): the reference's own in-core / direct runs take 10 / 11 iterations and end 1.5e-9 Eh apart; five runs of
    this test ended 1.5e-9 ... 1.3e-8 Eh from the anchors with 10 or 11 iterations.  What IS well defined is the
    converged state: with conver = 1e-11 the energy must equal the reference's converged value (its direct run stopped at
    RMS(P) = 2.8e-12) to 1e-9 Eh.  Round 1 allowed +12 iterations."""
    assert mol.is_converged and abs(mol.scf_iterations - a["iterations"]) <= 1, (name, mol.scf_iterations)
    assert abs(mol.energy.real - a["energy"]) < 5e-8, (name, mol.scf_iterations, mol.energy.real)


def test_ch4_converged_energy_matches_reference(golden):
    a = golden("anchors.json")["ch4_sto3g_direct"]          # P_RMS_final 2.8e-12: the reference's converged energy
    assert a["P_RMS_final"] < 1e-11
    for direct in (False, True):
        mol = Molecule(a["geometry"], a["basis"])
        mol.RHF(doPrint=False, direct=direct, conver=1e-11)
        assert mol.is_converged and abs(mol.energy.real - a["energy"]) < E_TOL, (direct, mol.energy.real)


@pytest.mark.parametrize("name", ["he_ccpvtz_incore", "h2co_sto3g_incore", "benzene_631gss_incore"])
def test_reference_smoke_configs_and_baseline_config2(golden, name):
    """He/cc-pVTZ (reference tests/test007.py), H2CO/STO-3G with its dipole (tests/test008.py) and benzene/6-31G**
    (BASELINE config 2: in-core tensor + one-pass J/K) against runs of the reference itself
    (tests/golden/make_golden_anchors2.py): identical iteration counts, 1e-9 Eh."""
    anchors = golden("anchors2.json")
    if name not in anchors:
        pytest.skip("anchor not generated yet (make_golden_anchors2.py --benzene)")
    a = anchors[name]
    mol = Molecule(a["geometry"], a["basis"])
    mol.RHF(doPrint=False, direct=a["direct"])
    assert mol.is_converged and mol.scf_iterations == a["iterations"]
    assert abs(mol.energy.real - a["energy"]) < E_TOL
    assert np.abs(np.real(np.asarray(mol.mu)) - np.asarray(a["dipole"])).max() < 1e-6


@pytest.mark.parametrize("tag,geom,basis", [("h2", "\n0 1\nH 0.0 0.0 0.74\nH 0.0 0.0 0.0\n", "sto-3g"), ("h2o", None, "sto-3g"),
                                            ("h2o_ccpvdz", None, "cc-pvdz")])
def test_nuclear_forces_vs_reference(golden, oracle, tag, geom, basis):
    """Molecule.forces() (device derivative integrals, csrc/grad.cu) against the forces the reference itself computed
    (tests/golden/grad_h2o.npz, written by mmd/forces.py of the reference with the P and F stored beside them): the
    gradient kernels are fed the reference's own converged P and F, so the comparison is at the 1e-10 level; the
    cc-pVDZ case runs d functions through f-type shifted primitives.  Also: the forces after our own SCF agree to
    the SCF convergence level, the literal of the reference's tests/test001.py, and zero net force."""
    g = golden("grad_h2o.npz")
    mol = Molecule(geom if geom is not None else synth.water(), basis)
    mol.RHF(doPrint=False)
    own = mol.forces().copy()
    ref = g["forces_" + tag]
    assert np.abs(own - ref).max() < 2e-7                      # densities converged to RMS(P) < 1e-8 on both sides
    assert np.abs(own.sum(axis=0)).max() < 1e-8                # translational invariance
    mol.P, mol.F = np.array(g["forces_" + tag + "_P"]), np.array(g["forces_" + tag + "_F"])
    got = mol.forces()
    assert np.abs(got - ref).max() < 1e-10, np.abs(got - ref).max()
    if tag == "h2":
        lit = np.array([[0.0, 0.0, -0.027679601], [0.0, 0.0, 0.027679601]])      # reference tests/test001.py:16-17
        assert np.abs(got - lit).max() < 5e-9
    # the two-electron part against the oracle's contraction (per canonical quartet, all four centres)
    masks = np.array([a.mask for a in mol.atoms])
    e2 = oracle.forces_2e_contracted(mol.bfs, masks, mol.P)
    assert np.abs(mol.gradient_parts["two_electron"] - e2).max() < 1e-10


def test_complex_density_updateFock_step(golden):
    """mol.updateFock() with a genuinely complex Hermitian density — what real-time propagation calls every step
    (reference mmd/realtime.py:62, mmd/scf.py:140-144) and the reason the J/K boundary is complex: P, J, K, F, FO
    against the reference's own arrays."""
    g = golden("updatefock_h2o.npz")
    mol = Molecule(synth.water(), "sto-3g")
    mol.RHF(doPrint=False)
    mol.PO = np.array(g["PO"])
    mol.updateFock()
    assert np.abs(mol.P - g["P"]).max() < 1e-12
    assert np.abs(mol.J - g["J"]).max() < FOCK_TOL and np.abs(mol.K - g["K"]).max() < FOCK_TOL
    assert np.abs(mol.F - g["F"]).max() < FOCK_TOL and np.abs(mol.FO - g["FO"]).max() < FOCK_TOL
    assert np.abs(mol.F.imag).max() > 1e-3          # the step really is complex


# ---- benchmark-size properties ---------------------------------------------------------------
def _core_guess_density(mol):
    import scipy.linalg
    mol.one_electron_integrals()
    FO = mol.X.T @ mol.Core @ mol.X
    _, CO = scipy.linalg.eigh(FO)
    C = mol.X @ CO
    occ = C[:, :mol.nocc]
    return (occ @ occ.conj().T).astype(complex)


@pytest.mark.parametrize("cfg", ["benzene_631gss", "w8_ccpvdz"])
def test_direct_build_equals_incore_build(cfg):
    """Every class (ss|ss)..(dd|dd), screening, digestion and the dense fill at configuration 2/3 size:
    sym(G_direct) with tol=0 must equal 2J-K from one pass over the dense tensor."""
    mol = Molecule(*synth.config(cfg))
    N = mol.nbasis
    P = _core_guess_density(mol)
    eng = mol.engine
    scr = eng.schwarz()
    G = eng.formPT(P, np.zeros_like(P), screen=scr, tol=0.0)
    G = 0.5 * (G + G.T)
    import torch
    n = eng.Ndev
    T = torch.empty((n, n, n, n), dtype=torch.float64, device=eng.tdev)
    from mmd._b200 import lib as L
    L.check(eng.lib.mmdb_eri_dense(eng.h, L.ptr(T), eng._stream()))
    # 8-fold symmetry of the device tensor
    assert torch.equal(T, T.permute(1, 0, 2, 3)) and torch.equal(T, T.permute(2, 3, 0, 1))
    J, K = eng.jk_incore(P, TwoE=T)
    assert np.abs(G - (2.0 * J - K)).max() < FOCK_TOL
    # default tolerance drops only negligible quartets
    G12 = eng.formPT(P, np.zeros_like(P), screen=scr, tol=1e-12)
    assert np.abs(0.5 * (G12 + G12.T) - G).max() < 1e-9
    del T


def closed_form_densities(bfs):
    """The two closed-form densities of tests/golden/make_golden_fock.py (bit-reproducible from the geometry)."""
    N = len(bfs)
    C = np.array([np.asarray(b.origin, dtype=np.float64) for b in bfs])
    i = np.arange(N, dtype=np.float64)
    A = 0.1 * np.cos(0.37 * (i[:, None] + i[None, :])) + 0.05 * np.cos(0.011 * (i[:, None] - i[None, :]) ** 2)
    A = 0.5 * (A + A.T)
    r2 = ((C[:, None, :] - C[None, :, :]) ** 2).sum(-1)
    B = np.exp(-0.35 * r2) * (0.3 * np.cos(0.61 * (i[:, None] + i[None, :])) + 0.2)
    B = 0.5 * (B + B.T)
    return {"A": A, "B": B}


@pytest.mark.parametrize("cfg", ["c20h42_631gs", "w32_ccpvdz"])
def test_benchmark_size_fock_elements_vs_oracle(golden, cfg):
    """Configurations 4 and 5 against the ORACLE (not against another GPU path): 43-48 stratified elements of
    sym(G) = 2J - K, each summed integral by integral (1e5-4e5 ERIs per element) by the pinned C oracle
    (tests/golden/make_golden_fock.py), for a dense and a local closed-form density.  Row chunking of the lists,
    the auxiliary stream, the screening pipeline and the bra slices only engage at these sizes; a chunk that was
    dropped consistently would show here.  tol = 0 (no screening) and the production tol = 1e-12."""
    g = golden("fock_elements_%s.npz" % cfg)
    mol = Molecule(*synth.config(cfg))
    assert mol.nbasis == int(g["nbasis"])
    eng = mol.engine
    scr = eng.schwarz()
    el = g["elements"]
    for name, P in closed_form_densities(mol.bfs).items():
        Pc = P.astype(complex)
        for tol in (0.0, 1e-12):
            G = eng.formPT(Pc, np.zeros_like(Pc), screen=scr, tol=tol)
            assert np.abs(G.imag).max() == 0.0
            Gs = 0.5 * (G.real + G.real.T)
            err = np.abs(Gs[el[:, 0], el[:, 1]] - g["G_" + name])
            assert err.max() < FOCK_TOL, (cfg, name, tol, el[err.argmax()], float(err.max()))


def test_full_size_direct_build_properties():
    """(H2O)_32/cc-pVDZ, 800 functions: shard additivity, linearity, stats consistency."""
    import ctypes as C
    import torch
    from mmd._b200 import lib as L
    mol = Molecule(*synth.config("w32_ccpvdz"))
    assert mol.nbasis == 800
    P = _core_guess_density(mol)
    eng = mol.engine
    scr = eng.schwarz()
    G = eng.formPT(P, np.zeros_like(P), screen=scr, tol=1e-12)
    st = eng.last_stats
    assert st["quartets"] > 1e8 and st["prim_quartets"] > st["quartets"]
    # shards: 4 partial builds into separate buffers must add up to the full build
    re = torch.from_numpy(np.ascontiguousarray(P.real)).to(eng.tdev)
    acc = torch.zeros((800, 800), dtype=torch.float64, device=eng.tdev)
    nq = 0
    for s in range(4):
        part = torch.zeros((800, 800), dtype=torch.float64, device=eng.tdev)
        stats = L.FockStats()
        L.check(eng.lib.mmdb_fock_direct(eng.h, L.ptr(re), None, 1e-12, L.ptr(part), None, s, 4, 0, C.byref(stats), eng._stream()))
        nq += stats.quartets
        acc += part
    assert nq == st["quartets"]
    assert np.abs(acc.cpu().numpy() - G.real).max() < FOCK_TOL
    # linearity in the density at tol = 0 (screening off)
    rng = np.random.default_rng(11)
    A = rng.standard_normal((800, 800)) * 1e-3
    P2 = (A + A.T).astype(complex)
    Ga = eng.formPT(P, np.zeros_like(P), screen=scr, tol=0.0)
    Gb = eng.formPT(P2, np.zeros_like(P), screen=scr, tol=0.0)
    Gab = eng.formPT(P + P2, np.zeros_like(P), screen=scr, tol=0.0)
    assert np.abs(Gab - Ga - Gb).max() < FOCK_TOL
    # incremental build: G(P) - G(P_old) == G(P - P_old)
    Ginc = eng.formPT(P + P2, P, screen=scr, tol=0.0)
    assert np.abs(Ginc - Gb).max() < FOCK_TOL
