"""Segment-sharded execution: one process (rank) per GPU, contiguous canonical segment slices
(SURVEY.md §8e).  The data path has no collective: every rank scans its own slice.  The only
exchange is one int64 per rank — its local match count (already capped at LIMIT).  On GPUs the
library does it itself, on the device (peer stores over NVLink behind the query's last kernel,
`SegmentManager.comm_connect`, include/imm3.h imm3_comm_*); without a connected communicator (the
gloo CPU tests, which exercise the host arithmetic) it is a `torch.distributed.all_gather`.  Either
way each rank then knows its global output offset and how many of its leading rows survive the LIMIT cut.

The local executor is any object with `begin(query) -> handle` (handle.local_count) and
`finish(handle, take) -> list of numpy columns`.  The product executor is `CudaExecutor`
(immutable3_b200.engine.Engine over the C ABI).  Tests inject other executors to exercise the
offset/LIMIT arithmetic without GPUs; nothing in this module computes query results itself.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .engine import Engine, Query, Result


def shard_range(nsegments: int, rank: int, world: int):
    """Canonical slice [rank*n/world, (rank+1)*n/world) — same arithmetic as the C library."""
    return rank * nsegments // world, (rank + 1) * nsegments // world


def limit_split(counts: Sequence[int], limit: int):
    """(offsets, takes): exclusive scan of the per-rank counts and the rows each rank emits so that
    the concatenation in rank order is the first `limit` rows of the global canonical order
    (limit <= 0: unlimited; Project.scala:73-77)."""
    offsets, takes, run = [], [], 0
    for c in counts:
        offsets.append(run)
        takes.append(c if limit <= 0 else max(0, min(c, limit - run)))
        run += c
    return offsets, takes


class CudaExecutor:
    def __init__(self, engine: Engine):
        self.engine = engine

    def begin(self, query: Query) -> Result:
        return self.engine.begin(query)

    def finish(self, handle: Result, take: int) -> List[np.ndarray]:
        handle.fetch(take)
        return handle.columns()


@dataclass
class ShardResult:
    rank: int
    world: int
    counts: List[int]      # local counts of every rank (each capped at LIMIT)
    offset: int            # global ordinal of this rank's first emitted row
    take: int              # rows this rank emits
    total: int             # rows of the whole result
    columns: List[np.ndarray]
    device_ms: float = 0.0


class ShardedEngine:
    def __init__(self, executor, group=None, collective_device: Optional[str] = None):
        import torch.distributed as dist

        self.executor = executor
        self.group = group
        self.dist = dist
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if collective_device is None:
            collective_device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        self.collective_device = collective_device

    def execute(self, query: Query) -> ShardResult:
        import torch

        h = self.executor.begin(query)
        if getattr(getattr(self.executor, "engine", None), "sm", None) is not None and self.executor.engine.sm.comm_connected:
            # the GPUs exchanged the counts themselves (imm3_comm_*): nothing to gather here
            counts = h.rank_counts
            cols = self.executor.finish(h, h.take)
            return ShardResult(self.rank, self.world, counts, h.global_offset, h.take, h.global_count, cols, float(h.device_ms))
        mine = torch.tensor([int(h.local_count)], dtype=torch.int64, device=self.collective_device)
        everyone = [torch.zeros_like(mine) for _ in range(self.world)]
        self.dist.all_gather(everyone, mine, group=self.group)  # the only exchange on the path
        counts = [int(t.item()) for t in everyone]
        offsets, takes = limit_split(counts, int(query.project.limit))
        cols = self.executor.finish(h, takes[self.rank])
        ms = float(getattr(h, "device_ms", 0.0))
        return ShardResult(self.rank, self.world, counts, offsets[self.rank], takes[self.rank], sum(takes), cols, ms)

    def gather_rows(self, res: ShardResult, dst: int = 0):
        """Concatenate every rank's emitted rows on `dst` in rank order (= canonical order). Test helper."""
        payload = [c for c in res.columns]
        gathered = [None] * self.world if self.rank == dst else None
        self.dist.gather_object(payload, gathered, dst=dst, group=self.group)
        if self.rank != dst:
            return None
        ncols = len(payload)
        return [np.concatenate([g[c] for g in gathered]) for c in range(ncols)]
