"""Multi-GPU plumbing: one process per GPU (torch.distributed), seed-cell units sharded rank-strided, ONE reduce
(sum, int64) of the count tensor + division counters at the end of a run.  The reference has no multi-GPU path
(device 0 is hard-coded: src/simulation/proliferation.cu:38, cells_population.cu:32)."""
from __future__ import annotations

import torch
import torch.distributed as dist

DEFAULT_SHARD_UNIT = 32


def shard_spec(rank: int | None = None, world: int | None = None, unit: int = DEFAULT_SHARD_UNIT):
    """(rank, world, unit) for procell_sim_params: this rank simulates the seed-cell units u with u % world == rank."""
    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
        else:
            rank, world = 0, 1
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world %d" % (rank, world))
    return (rank, world, unit)


def owner_of_seed(root: int, world: int, unit: int = DEFAULT_SHARD_UNIT) -> int:
    return (root // unit) % world


def owner_of_shared_node(root: int, world: int) -> int:
    """Subtree sharding (procell_sim_params.shard_level = L >= 1): every rank expands the nodes of tree level < L; what
    they count (level-0 leaves, their divisions, the leaves among their daughters) is credited to this rank."""
    return root % world


def owner_of_subtree(root: int, heap: int, world: int) -> int:
    """... and a daughter at level L (heap index in [2^L, 2^(L+1))) that will divide is kept by this rank alone."""
    return ((root + (heap & 0xFFFFFFFF)) & 0xFFFFFFFF) % world


def packed_buffer(n_sets: int, n_keys: int, n_types: int, device) -> torch.Tensor:
    """counts [n_sets*n_keys*n_types] followed by divisions [n_sets]: one tensor, so one collective."""
    return torch.zeros(n_sets * n_keys * n_types + n_sets, dtype=torch.int64, device=device)


def reduce_packed(buf: torch.Tensor, dst: int = 0) -> torch.Tensor:
    """The single exchange step of the path: sum the per-GPU count tensors onto `dst` (NCCL over NVLink on GPUs,
    gloo in the CPU tests).  No-op for a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM)
    return buf


def unpack(buf: torch.Tensor, n_sets: int, n_keys: int, n_types: int):
    n = n_sets * n_keys * n_types
    return buf[:n].view(n_sets, n_keys, n_types), buf[n:n + n_sets]
