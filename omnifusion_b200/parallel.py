"""Panorama-level data parallelism: one process per GPU, the batch split into contiguous equal
shards, no data-path collective (every panorama is independent in eval mode); the only exchange
is the final all-reduce of metric partial sums.  Replaces the reference's single-process
nn.DataParallel scatter/gather (test.py:107)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun). Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            # NCCL_DEBUG is left exactly as the launcher set it (its log is how the rank count is verified);
            # bench.py keeps its single JSON line clean by diverting fd 1 itself
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_bounds(total, rank, world):
    """Contiguous shard [lo, hi) of `total` panoramas for `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch, rank, world):
    lo, hi = shard_bounds(batch.shape[0], rank, world)
    return batch[lo:hi]


def reduce_sums(partials):
    """all-reduce(SUM) of a small float64 tensor of partial sums (e.g. 7 metrics x (sum, n))."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(partials, op=dist.ReduceOp.SUM)
    return partials


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            dist.barrier(device_ids=[torch.cuda.current_device()])
        else:
            dist.barrier()
