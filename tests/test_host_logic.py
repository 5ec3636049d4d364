"""Host-side logic of the drop-in package, no GPU: basis tables, Basis normalisation, shell
reconstruction, C-ABI library loading / exported symbols, loud failure without a device, and the
Molecule / SCF / PostSCF drivers exercised on an oracle-backed engine (tests/oracle_engine.py)."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle_engine
from conftest import ROOT
from mmd._b200 import basisio, dist, lib, synth
from mmd._b200.shells import ShellTable, cart_components
from mmd.integrals.twoe import Basis


def _no_gpu():
    try:
        lib.require_gpu()
        return False
    except lib.MMDBError:
        return True


# ---- C ABI ------------------------------------------------------------------------------------
def test_library_loads_and_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mmdb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mmdb_[A-Za-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    L = ctypes.CDLL(lib.SO_PATH)
    for name in declared:
        assert hasattr(L, name), "missing export " + name
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)
    assert lib.load().mmdb_version() >= 100


def test_stats_struct_layout_matches_the_header(tmp_path):
    """The ctypes mirror of mmdb_fock_stats (mmd/_b200/lib.py) has the size and field offsets the C compiler gives the
    struct of include/mmdb200.h (a field added on one side only would shift every counter read through the C ABI)."""
    import ctypes as C
    import subprocess
    from mmd._b200 import lib as L
    src = tmp_path / "layout.c"
    fields = [f for f, _ in L.FockStats._fields_]
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mmdb200.h"\nint main(void){printf("%zu", sizeof(mmdb_fock_stats));'
                   + "".join('printf(" %%zu", offsetof(mmdb_fock_stats, %s));' % f for f in fields) + "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert got[0] == C.sizeof(L.FockStats)
    assert got[1:] == [getattr(L.FockStats, f).offset for f in fields]


def test_flop_model_matches_survey_table():
    # SURVEY.md §8(d) per-class values
    table = {(0, 0, 0, 0): 59, (1, 0, 0, 0): 101, (1, 0, 1, 0): 213, (1, 1, 0, 0): 259, (1, 1, 1, 0): 581,
             (1, 1, 1, 1): 1732, (2, 0, 0, 0): 208, (2, 1, 1, 0): 1435, (2, 1, 1, 1): 4143, (2, 2, 0, 0): 1717,
             (2, 1, 2, 1): 9653, (2, 2, 1, 1): 10109, (2, 2, 2, 1): 22243, (2, 2, 2, 2): 51900}
    for cls, f in table.items():
        assert lib.class_flops(*cls) == f


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_device():
    from mmd._b200 import engine
    from mmd.molecule import Molecule
    mol = Molecule(synth.water(), "sto-3g")
    with pytest.raises(lib.MMDBError):
        engine.Engine(mol.bfs)
    with pytest.raises(lib.MMDBError):
        mol.RHF(doPrint=False)


# ---- basis data -------------------------------------------------------------------------------
def test_basis_tables():
    d = basisio.load_basis("cc-pvdz", os.path.join(ROOT, "mcmurchie-davidson_b200", "mmd", "basis"))
    assert [(m, len(p)) for m, p in d[8]] == [("S", 8), ("S", 8), ("S", 1), ("P", 3), ("P", 1), ("D", 1)]
    assert [(m, len(p)) for m, p in d[1]] == [("S", 3), ("S", 1), ("P", 1)]
    d = basisio.load_basis("6-31gss", os.path.join(ROOT, "mcmurchie-davidson_b200", "mmd", "basis"))
    assert [(m, len(p)) for m, p in d[6]] == [("S", 6), ("S", 3), ("P", 3), ("S", 1), ("P", 1), ("D", 1)]
    assert d[6][1][1][0][0] == d[6][2][1][0][0]          # SP shell: shared exponents


def test_g94_reader_roundtrip(tmp_path):
    txt = "!comment\n****\nH 0\nS 2 1.00\n 1.0D+01 0.5\n 2.0 0.25\nSP 1 1.00\n 0.5 1.0 0.7\n****\n"
    p = tmp_path / "x.gbs"
    p.write_text(txt)
    d = basisio.parse_g94(str(p))
    assert d == {1: [("S", [(10.0, 0.5), (2.0, 0.25)]), ("S", [(0.5, 1.0)]), ("P", [(0.5, 0.7)])]}


def test_basis_sizes_of_benchmark_configs():
    from mmd.molecule import Molecule
    sizes = {"h2o_sto3g": 7, "h2o_ccpvdz": 25, "benzene_631gss": 120, "w8_ccpvdz": 200, "c20h42_631gs": 384, "w32_ccpvdz": 800}
    for cfg, n in sizes.items():
        assert Molecule(*synth.config(cfg)).nbasis == n


def test_basis_normalisation_vs_oracle(oracle):
    e = [3047.5249, 457.36951, 103.94869, 29.210155, 9.286663, 3.163927]
    c = [0.0018347, 0.0140373, 0.0688426, 0.2321844, 0.4679413, 0.3623120]
    for lmn in [(0, 0, 0), (1, 0, 0), (0, 0, 1), (2, 0, 0), (1, 1, 0), (0, 1, 1), (0, 0, 2)]:
        b = Basis([0.1, 0.2, 0.3], lmn, len(e), e, c)
        cc, nn = oracle.normalize(lmn, e, c)
        assert np.allclose(b.coefs, cc, rtol=1e-14, atol=0) and np.allclose(b.norm, nn, rtol=1e-14, atol=0)
        assert b.shell.dtype == np.int64 and b.origin.shape == (3,) and b.num_exps == 6


def test_shell_table_full_and_ghost_shells():
    from mmd.molecule import Molecule
    mol = Molecule(synth.water(), "cc-pvdz")
    t = ShellTable(mol.bfs)
    assert t.identity and t.nshell == 12 and t.ndev == 25
    assert t.am.tolist() == [0, 0, 0, 1, 1, 2, 0, 0, 1, 0, 0, 1]
    # hand-built list: a lone d_xy, an s, a full p shell in order, a lone p_z
    mk = lambda lmn, x: Basis([x, 0, 0], lmn, 1, [0.8], [1.0])
    bfs = [mk((1, 1, 0), 0.0), mk((0, 0, 0), 0.0), mk((1, 0, 0), 1.0), mk((0, 1, 0), 1.0), mk((0, 0, 1), 1.0), mk((0, 0, 1), 2.0)]
    t = ShellTable(bfs)
    assert not t.identity
    assert t.am.tolist() == [2, 0, 1, 1] and t.ndev == 6 + 1 + 3 + 3
    assert t.user2dev.tolist() == [1, 6, 7, 8, 9, 12]
    M = np.arange(36.0).reshape(6, 6)
    assert np.array_equal(t.to_user_matrix(t.to_dev_matrix(M)), M)
    assert cart_components(2) == [(2, 0, 0), (1, 1, 0), (1, 0, 1), (0, 2, 0), (0, 1, 1), (0, 0, 2)]


def test_formPT_rejects_real_density():
    from mmd.integrals.fock import formPT
    with pytest.raises(ValueError):
        formPT(np.zeros((2, 2)), np.zeros((2, 2)), [None, None], 2, {}, 1e-12)


def test_static_shard_schedule_is_balanced():
    # rows sorted by contraction depth (as the library orders them): round-robin dealing balances cost
    rng = np.random.default_rng(0)
    cost = np.sort(rng.integers(1, 65, size=5000))[::-1].astype(float) ** 2
    for n in (2, 4, 8):
        c = dist.shard_costs(cost, n)
        assert c.max() / c.mean() < 1.02
        owned = np.concatenate([dist.shard_rows(len(cost), s, n) for s in range(n)])
        assert np.array_equal(np.sort(owned), np.arange(len(cost)))


# ---- drivers on the oracle-backed engine ------------------------------------------------------
ANCHORS_EXACT = ["h2_sto3g_incore", "h2o_sto3g_incore", "h2o_sto3g_direct", "ch4_321g_incore", "he2_ccpvdz_incore",
                 "h2o_dz_incore", "h2o_321g_incore"]


@pytest.mark.parametrize("name", ANCHORS_EXACT)
def test_scf_driver_reproduces_reference_trajectory(monkeypatch, golden, name):
    oracle_engine.install(monkeypatch)
    from mmd.molecule import Molecule
    from mmd.postscf import PostSCF
    a = golden("anchors.json")[name]
    geom = a.get("geometry", synth.water())
    mol = Molecule(geom, a.get("basis", "sto-3g"))
    mol.RHF(doPrint=False, direct=a.get("direct", False), conver=a.get("conver", 1e-8))
    assert mol.is_converged
    assert mol.scf_iterations == a["iterations"]                      # identical iteration count
    assert abs(mol.energy.real - a["energy"]) < 1e-9                   # Eh
    assert np.abs(np.array([h[0] for h in mol.scf_history]) - np.array(a["energies"])).max() < 1e-8
    if "emp2" in a:
        PostSCF(mol).MP2()
        assert abs(mol.emp2.real - a["emp2"]) < 1e-9


def test_scf_degenerate_guess_case_is_noise_limited(monkeypatch, golden):
    """CH4/STO-3G: the core-Hamiltonian guess splits a degenerate t2 set across the occupied/virtual
    boundary, so the trajectory depends on rounding noise; the reference's own in-core and direct
    runs differ (10 vs 11 iterations).  Only the converged energy is comparable,
    to the ~1e-8 the default convergence threshold leaves."""
    oracle_engine.install(monkeypatch)
    from mmd.molecule import Molecule
    for name in ("ch4_sto3g_incore", "ch4_sto3g_direct"):
        a = golden("anchors.json")[name]
        mol = Molecule(a["geometry"], a["basis"])
        mol.RHF(doPrint=False, direct=a["direct"])
        assert mol.is_converged and abs(mol.scf_iterations - a["iterations"]) <= 1
        assert abs(mol.energy.real - a["energy"]) < 5e-8


def test_tight_convergence_is_noise_limited(monkeypatch, golden):
    """conver=1e-14 is below the rounding noise of the Fock build: the iteration count depends on
    last-bit differences (the reference needs 20 iterations, the oracle-backed driver 14); only
    convergence and the energy are comparable."""
    oracle_engine.install(monkeypatch)
    from mmd.molecule import Molecule
    a = golden("anchors.json")["h2o_sto3g_incore_tight"]
    mol = Molecule(synth.water(), "sto-3g")
    mol.RHF(doPrint=False, conver=1e-14)
    assert mol.is_converged and abs(mol.energy.real - a["energy"]) < 1e-9


@pytest.mark.parametrize("direct", [False, True])
def test_device_resident_scf_loop_matches_host_loop(monkeypatch, direct):
    """mmd/scf.py:_RHF_device (torch tensors, eigh/DIIS/energy on the tensor device) against the NumPy/SciPy loop,
    both fed by the oracle: same trajectory, iteration count, energy and final matrices (CPU tensors here; the
    GPU suite repeats it on the device)."""
    import oracle_engine
    from mmd._b200 import synth
    from mmd.molecule import Molecule
    oracle_engine.install(monkeypatch)
    host = Molecule(synth.water(), "sto-3g")
    host.RHF(doPrint=False, direct=direct)
    oracle_engine.install(monkeypatch, oracle_engine.OracleTensorEngine)
    dev = Molecule(synth.water(), "sto-3g")
    dev.RHF(doPrint=False, direct=direct)
    assert dev.is_converged and host.is_converged and dev.scf_iterations == host.scf_iterations
    assert abs(dev.energy.real - host.energy.real) < 1e-10
    assert len(dev.scf_history) == len(host.scf_history)
    assert max(abs(a[0] - b[0]) for a, b in zip(dev.scf_history, host.scf_history)) < 1e-9
    for attr in ("P", "F", "MO", "G", "P_old"):
        assert isinstance(getattr(dev, attr), np.ndarray)
        assert np.abs(getattr(dev, attr) - getattr(host, attr)).max() < 1e-8, attr
    assert np.abs(np.asarray(dev.mu) - np.asarray(host.mu)).max() < 1e-7


def test_forces_need_a_converged_scf():
    """Molecule.forces() mirrors the reference's guard (mmd/forces.py:11-12): no gradient before the SCF has converged."""
    from mmd.molecule import Molecule
    from mmd._b200 import synth
    mol = Molecule.__new__(Molecule)
    mol.is_converged = False
    with pytest.raises(SystemExit):
        mol.forces()


def test_printed_summary_format(monkeypatch, capsys):
    oracle_engine.install(monkeypatch)
    from mmd.molecule import Molecule
    from mmd.postscf import PostSCF
    mol = Molecule(synth.water(), "sto-3g")
    mol.RHF()
    PostSCF(mol).MP2()
    out = capsys.readouterr().out
    assert re.search(r"E\(SCF\)    =  -74\.9420798\d+ in 9 iterations", out)
    assert "FPS-SPF" in out and "RMS(P)" in out and "Dipole Y =  1.534009" in out
    assert re.search(r"E\(MP2\) =  -74\.99122954", out)
    assert mol.TwoE.shape == (7, 7, 7, 7) and isinstance(mol.TwoE, np.ndarray)


def test_geometry_units_and_center_of_charge():
    from mmd.molecule import Molecule
    mol = Molecule(synth.water(), "sto-3g")
    assert abs(mol.atoms[1].origin[0] - 0.866811829 / 0.52917721092) < 1e-15
    assert mol.nelec == 10 and mol.nocc == 5 and mol.charge == 0 and mol.multiplicity == 1
    z = np.array([8, 1, 1.0])
    xyz = np.array([a.origin for a in mol.atoms])
    assert np.allclose(mol.center_of_charge, (z[:, None] * xyz).sum(0) / z.sum(), atol=1e-15)


def test_bench_reference_arm_contract():
    """bench.py --impl reference: one JSON line with the contract's keys, timed on the host with oracle/_ref (or
    the oracle port), no GPU involved."""
    import json
    import subprocess
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    if "unavailable" in line:
        return
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["metric"] == "screened_eri_shell_quartets_per_s_direct_fock_build" and line["unit"] == "quartets/s"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
