"""Multi-GPU plumbing for the direct Fock build: one process per GPU (torch.distributed),
static sharding of KET shell-pair rows, one sum all-reduce of the partial G matrices.

The shard rule is the one the screening kernel applies (csrc/lib.cu screen_kernel): ket pair row j of
every class pair belongs to shard j % nshards (the columns of a row are the bra pairs).  Rows are ordered by contraction depth inside a
class, so dealing them round-robin gives every shard the same mix of cheap and expensive rows
(static, cost-balanced, no communication on the data path)."""
import numpy as np


def world():
    """(rank, world_size) of the default process group, (0, 1) when none is initialised."""
    try:
        import torch.distributed as dist
    except Exception:
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_of_row(i, nshards):
    return np.asarray(i) % int(nshards)


def shard_rows(nrows, shard, nshards):
    """Rows of a pair class owned by `shard`."""
    return np.arange(int(shard), int(nrows), int(nshards))


def shard_costs(row_cost, nshards):
    """Total model cost per shard for per-row costs (diagnostic: balance of the static schedule)."""
    row_cost = np.asarray(row_cost, dtype=np.float64)
    return np.bincount(np.arange(len(row_cost)) % nshards, weights=row_cost, minlength=nshards)


_COMM = {}


def c_abi_comm(lib, device):
    """The library's own NCCL communicator (mmdb_comm_init, include/mmdb200.h) over the ranks of the default
    torch.distributed group: torch is only the rendezvous that carries the 128-byte unique id from rank 0."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    rank, size = world()
    key = (int(device), rank, size)
    if key in _COMM:
        return _COMM[key]
    from . import lib as L
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        L.check(lib.mmdb_comm_unique_id(buf))
        uid = torch.tensor(list(buf), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = uid.to(torch.device("cuda", int(device)))
        dist.broadcast(t, 0)
        uid = t.cpu()
    else:
        dist.broadcast(uid, 0)
    raw = (C.c_ubyte * 128)(*uid.tolist())
    h = C.c_void_p()
    L.check(lib.mmdb_comm_init(int(device), size, rank, raw, C.byref(h)))
    _COMM[key] = h
    return h


def allreduce_sum_(tensor):
    """In-place sum over ranks (NCCL on GPUs, gloo in the CPU tests); no-op without a group."""
    rank, size = world()
    if size > 1:
        import torch.distributed as dist
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor
