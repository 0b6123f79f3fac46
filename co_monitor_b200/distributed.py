"""Multi-GPU plumbing: sharding helpers and the single reduce of the path.

Stage 1 is independent per voxel and Stage 2 is a sum over voxels, so the path shards by voxels with *no* data-path
collective in Stage 1.  The production shards are block-cyclic over the lattice's z-columns (``Lattice.shard``,
``ComLattice.shard_*``: the visible voxels are spatially clustered, contiguous ranges do not balance);
``shard_range`` is the contiguous split, used for explicit voxel arrays and in tests.  The only exchange is one
all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests) of a packed fp64 buffer - the Reff-millimetre
weight histograms from which every rank then evaluates Stage 2, weight sums - and of the int64 counters.  Per-voxel
weight arrays stay sharded on their ranks.  tests/test_distributed_gloo.py runs both forms on two CPU ranks.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of rank *rank* out of *world* (sizes differ by at most one voxel)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return rank * n // world, (rank + 1) * n // world


@dataclass
class PackedResults:
    """Named fp64 arrays + int64 counters that travel in one all-reduce each."""

    floats: dict = field(default_factory=dict)
    counters: dict = field(default_factory=dict)

    def add(self, name: str, value) -> None:
        self.floats[name] = value

    def count(self, name: str, value: int) -> None:
        self.counters[name] = int(value)

    def all_reduce(self, group=None, device=None) -> "PackedResults":
        """Sum over ranks (in place); returns self.  Works on CPU tensors (gloo) and CUDA tensors (nccl)."""
        import torch
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return self
        names = sorted(self.floats)
        tensors = [torch.as_tensor(self.floats[k], dtype=torch.float64, device=device).reshape(-1) for k in names]
        sizes = [t.numel() for t in tensors]
        if tensors:
            flat = torch.cat(tensors) if len(tensors) > 1 else tensors[0].clone()
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            off = 0
            for k, size in zip(names, sizes):
                orig = self.floats[k]
                piece = flat[off:off + size]
                off += size
                if torch.is_tensor(orig):
                    self.floats[k] = piece.reshape(orig.shape)
                else:
                    self.floats[k] = piece.cpu().numpy().reshape(np.shape(orig))
        cnames = sorted(self.counters)
        if cnames:
            c = torch.tensor([self.counters[k] for k in cnames], dtype=torch.int64, device=device)
            dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
            for k, v in zip(cnames, c.cpu().tolist()):
                self.counters[k] = int(v)
        return self
