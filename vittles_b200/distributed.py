"""Observation sharding helpers.

One process per GPU; observations (rows of X, hence columns of the sensitivity
matrix) are split contiguously; the only data-path collective is the all-reduce
of D-sized quantities (Hessian D x D, gradient, Hessian-vector products)."""
import os

import torch
import torch.distributed as dist


def shard_range(n_total, rank, world_size):
    """Contiguous row range [r0, r1) of `rank`; sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError('rank {} outside world of size {}'.format(rank, world_size))
    return (n_total * rank) // world_size, (n_total * (rank + 1)) // world_size


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK,
    LOCAL_RANK, WORLD_SIZE, MASTER_ADDR/PORT).  Returns (rank, world, group);
    group is None in a single-process run."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if world == 1:
        return rank, world, None
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if not dist.is_initialized():
        dist.init_process_group(backend)
    return rank, world, dist.group.WORLD


def allreduce_sum_(t, group=None):
    """In-place sum over the group (no-op without a group)."""
    if group is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def gather_columns(local_cols, group=None):
    """Concatenate per-rank column blocks of a (D, N_local) matrix along the
    observation axis (tests and small problems only)."""
    if group is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_cols
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=local_cols.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local_cols.shape[1]], dtype=torch.int64, device=local_cols.device),
                    group=group)
    # all_gather needs equal shapes: pad every block to the widest, trim afterwards
    widest = max(int(s.item()) for s in sizes)
    rows = local_cols.shape[0]
    padded = torch.zeros((rows, widest), dtype=local_cols.dtype, device=local_cols.device)
    padded[:, :local_cols.shape[1]] = local_cols
    outs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(outs, padded, group=group)
    return torch.cat([o[:, :int(s.item())] for o, s in zip(outs, sizes)], dim=1)
