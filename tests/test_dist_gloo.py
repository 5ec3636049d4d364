"""world_size-2 `gloo` test of the multi-GPU plumbing (CPU): every rank builds the partial G of the
quartets whose bra row it owns under the static shard rule, the partial matrices are summed with the
same all-reduce helper the GPU path uses, and the total must equal the unsharded oracle build."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT  # noqa: F401  (sets sys.path)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from mmd._b200 import dist as D, synth
    from mmd.molecule import Molecule
    assert D.world() == (rank, world)
    mol = Molecule(synth.water(), "sto-3g")
    N = mol.nbasis
    fb = O.FlatBasis(mol.bfs)
    scr = O.schwarz(fb)
    rng = np.random.default_rng(5)
    A = rng.standard_normal((N, N))
    P = (0.1 * (A + A.T)).astype(complex)
    # canonical quartets, bra row = pair index ij; the owner of a row is ij % world (the kernel's rule
    # applied to function pairs here — the host-side property under test is coverage + the reduction)
    G = np.zeros((N, N), dtype=complex)
    idx, meta = [], []
    for i in range(N):
        for j in range(i + 1):
            ij = i * (i + 1) // 2 + j
            if D.shard_of_row(ij, world) != rank:
                continue
            for k in range(N):
                for l in range(k + 1):
                    if ij >= k * (k + 1) // 2 + l:
                        idx.append((i, j, k, l))
    vals = O.ERI_batch(fb, np.array(idx))
    for (i, j, k, l), v in zip(idx, vals):
        deg = (1.0 if i == j else 2.0) * (1.0 if k == l else 2.0) * (1.0 if (i == k and j == l) else 2.0)
        e = deg * v
        G[i, j] += P[k, l] * e; G[k, l] += P[i, j] * e
        G[i, k] -= 0.25 * P[j, l] * e; G[j, l] -= 0.25 * P[i, k] * e
        G[i, l] -= 0.25 * P[j, k] * e; G[k, j] -= 0.25 * P[i, l] * e
    t = torch.from_numpy(np.stack([G.real, G.imag]))
    D.allreduce_sum_(t)
    total = t[0].numpy() + 1j * t[1].numpy()
    ref = O.formPT(P, np.zeros_like(P), fb, N, scr, 0.0)
    np.save(os.path.join(out_dir, "err_%d.npy" % rank), np.array([np.abs(total - ref).max(), len(idx)]))
    dist.destroy_process_group()


def test_sharded_build_sums_to_full_build(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    counts = 0
    for r in range(world):
        err, n = np.load(tmp_path / ("err_%d.npy" % r))
        assert err < 1e-12
        counts += int(n)
    assert counts == 406            # every canonical quartet of H2O/STO-3G owned by exactly one rank
