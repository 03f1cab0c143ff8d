"""
Data-parallel plumbing (one process per GPU, torchrun-style env): contiguous sharding of a batch of
sequences (training) or a set of songs (sampling) across ranks, and process-group set-up.
Training: every rank steps its slice with means taken over the GLOBAL batch (clv_cfg.B_global), the
flat [grads | loss scalars] buffer is sum-all-reduced over NCCL/NVLink, Adam-WN runs redundantly on
every rank.  Sampling: songs are independent, no collective; noise is keyed by the global song index
so the result does not depend on the number of ranks.
"""
import os
import torch
import torch.distributed as dist


def shard_range(n, world_size, rank):
    """Contiguous [lo, hi) slice of n items for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """torchrun env (RANK, LOCAL_RANK, WORLD_SIZE, MASTER_*) -> (world_size, rank, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return world, rank, local


def allreduce_sum_(flat, group=None):
    """Sum-all-reduce of the flat [grads | 8 loss scalars] buffer (in place)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def local_batch(order, step, B_local, world_size, rank):
    """Indices of the sequences rank `rank` steps at optimisation step `step` of an epoch whose GLOBAL sample
    order is `order` (identical on every rank): global batch `step` = order[step*Bg : (step+1)*Bg] with
    Bg = B_local * world_size, of which each rank takes its contiguous B_local slice -- so N ranks with
    --batch_size Bg visit exactly the batches one process with --batch_size Bg would."""
    Bg = B_local * world_size
    lo = step * Bg + rank * B_local
    return order[lo:lo + B_local]


def shared_permutation(n, shuffle, device=None, group=None):
    """The epoch's sample order, drawn on rank 0 (np.random, like Keras' fit) and broadcast to every rank."""
    import numpy as np
    perm = torch.from_numpy(np.random.permutation(n) if shuffle else np.arange(n))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        perm = perm.to(device) if device is not None else perm
        dist.broadcast(perm, src=0, group=group)
    return perm
