"""Data-parallel plumbing: one process per GPU, torch.distributed over NCCL/NVLink.

The reference has no multi-device code at all; this is new functionality.
Training batches are sharded by sentence across ranks; each step exchanges ONE buffer
(flat gradients + loss sum + token count, `ParameterArena.allreduce_view`) with one
all-reduce(sum); parameters and optimizer state are replicated.  Beam search stays
single-GPU (BASELINE.json north_star).
"""
import os
from typing import List, Sequence

import torch
import torch.distributed as dist


def is_initialized() -> bool:
    return dist.is_available() and dist.is_initialized()


def init_from_env(backend: str = None) -> None:
    """Initialise the default process group from torchrun's environment (no-op for 1 rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or is_initialized():
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    dist.init_process_group(backend=backend)
    import atexit
    atexit.register(shutdown)


_cleanups = []    # callables run by shutdown() before the process group goes away


def register_cleanup(fn) -> None:
    """Something that must be released before the communicator is destroyed - a trainer's CUDA graphs that
    captured collectives: tearing NCCL down under them hangs the process at exit."""
    if fn not in _cleanups:
        _cleanups.append(fn)


def shutdown() -> None:
    """Release what was registered, drain the device, destroy the process group.  Registered with atexit by
    `init_from_env`; idempotent."""
    while _cleanups:
        try:
            _cleanups.pop()()
        except Exception:  # pylint: disable=broad-except
            pass
    if is_initialized():
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        dist.destroy_process_group()


def world_size() -> int:
    return dist.get_world_size() if is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if is_initialized() else 0


def all_reduce_sum(buf: torch.Tensor) -> None:
    """In-place sum over ranks (K14).  Asynchronous w.r.t. the host on NCCL."""
    if world_size() > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)


def all_reduce_async(buf: torch.Tensor):
    """Start an in-place sum over ranks and return its handle (None for one rank).  On NCCL the collective
    runs on the process group's own stream, ordered after everything already issued on the current stream;
    `handle.wait()` makes the current stream wait for it."""
    if world_size() > 1:
        return dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=True)
    return None


def barrier() -> None:
    """All ranks wait here (no-op for one rank): e.g. before reading a checkpoint rank 0 has just written."""
    if world_size() > 1:
        dist.barrier()


def shard_bounds(n_items: int, n_ranks: int) -> List[int]:
    """Contiguous split of n_items sentences into n_ranks shards, sizes differing by <= 1."""
    base, extra = divmod(n_items, n_ranks)
    bounds = [0]
    for r in range(n_ranks):
        bounds.append(bounds[-1] + base + (1 if r < extra else 0))
    return bounds


def shard(items: Sequence, r: int = None, n: int = None) -> Sequence:
    r = rank() if r is None else r
    n = world_size() if n is None else n
    b = shard_bounds(len(items), n)
    return items[b[r]:b[r + 1]]
