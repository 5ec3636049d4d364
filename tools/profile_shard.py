#!/usr/bin/env python
"""One shard of an N-way sharded direct Fock build on ONE GPU (no collective): per-class times of
shard 0 of `nshards`, and the un-instrumented wall time of every shard count in the list.
    python tools/profile_shard.py w32_ccpvdz 1,2,4,8
Shows which classes stop scaling when the ket rows are dealt out to more GPUs."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "mcmurchie-davidson_b200"))
from mmd._b200 import lib as L, synth   # noqa: E402
from mmd.molecule import Molecule        # noqa: E402


def main(workload, shard_counts):
    import scipy.linalg
    import torch
    mol = Molecule(*synth.config(workload))
    mol.one_electron_integrals()
    FO = mol.X.T @ mol.Core @ mol.X
    _, CO = scipy.linalg.eigh(FO)
    Cm = mol.X @ CO
    P = (Cm[:, :mol.nocc] @ Cm[:, :mol.nocc].conj().T).astype(complex)
    eng = mol.engine
    eng.schwarz()
    N = mol.nbasis
    dev = eng.tdev
    dP = torch.from_numpy(np.ascontiguousarray(P.real)).to(dev)
    G = torch.zeros((N, N), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev)

    def build(shard, nshards, stats=None, flags=0):
        G.zero_()
        L.check(eng.lib.mmdb_fock_direct(eng.h, L.ptr(dP), None, 1e-12, L.ptr(G), None, shard, nshards, flags,
                                         C.byref(stats) if stats is not None else None, C.c_void_p(stream.cuda_stream)))

    tables = {}
    for ns in shard_counts:
        for _ in range(3):
            build(0, ns)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record(stream)
        for _ in range(reps):
            build(0, ns)
        e1.record(stream)
        torch.cuda.synchronize()
        wall = e0.elapsed_time(e1) / reps
        st = L.FockStats()
        build(0, ns, st, flags=1)
        torch.cuda.synchronize()
        d = st.as_dict()
        tables[ns] = d
        tot = sum(v["ms"] + v["screen_ms"] for v in d["classes"].values())
        print("nshards %d: shard 0 wall %.3f ms (ideal %.3f), quartets %d, serialized class sum %.3f ms"
              % (ns, wall, 0.0 if ns == shard_counts[0] else tables[shard_counts[0]]["wall"] * shard_counts[0] / ns, d["quartets"], tot))
        d["wall"] = wall
    base = tables[shard_counts[0]]
    print("%-8s" % "class" + "".join("   eri/scr ms @%d" % ns for ns in shard_counts))
    for name in base["classes"]:
        row = "%-8s" % name
        for ns in shard_counts:
            c = tables[ns]["classes"].get(name)
            row += "   %6.3f/%6.3f" % ((c["ms"], c["screen_ms"]) if c else (0.0, 0.0))
        print(row)


if __name__ == "__main__":
    wl = sys.argv[1] if len(sys.argv) > 1 else "w32_ccpvdz"
    counts = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,8").split(",")]
    main(wl, counts)
