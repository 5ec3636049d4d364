"""Multi-GPU plumbing for the direct Fock build: one process per GPU (torch.distributed),
static sharding of bra shell-pair rows, one sum all-reduce of the partial G matrices.

The shard rule is the one the screening kernel applies (csrc/lib.cu screen_kernel): bra pair row i of
every pair class belongs to shard i % nshards.  Rows are ordered by contraction depth inside a
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


def allreduce_sum_(tensor):
    """In-place sum over ranks (NCCL on GPUs, gloo in the CPU tests); no-op without a group."""
    rank, size = world()
    if size > 1:
        import torch.distributed as dist
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor
