"""Sample / component sharding over one process per GPU (torch.distributed, NCCL over NVLink; gloo in CPU tests).

Partitioning: the iteration's global sample index space [0, N) in the reference's order
(component-major, models/gmm.py:378-386) is cut into `world` contiguous ranges; rank r draws the noise of ITS rows
from the counter-based generator (value depends only on the global row), evaluates log-densities / target /
gradients for them, and contributes partial per-component sums.  Exchange steps:
  * importance-weight normalisers: all-reduce MAX [K] + 2 x all-reduce SUM [K]
  * Stein statistics (-E[H]) [K,D,D] and (-E[g]) [K,D]: reduce-scatter by component when K divides evenly (the
    component update only needs its own shard), all-reduce otherwise
  * expected log-ratios for the weight update: all-reduce SUM [K]
  * component update: components are sharded K/world per rank, the new (mean, Cholesky) are all-gathered; the
    gathers of the Cholesky factors and of the precisions run asynchronously and are only waited for at their
    first use (next sampling step / next gradient), overlapping the weight-update log-density pass.
Model parameters and all O(K) learner state are replicated and stay bit-identical on every rank because every
collective returns the same bits to all ranks."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


class ShardContext:
    def __init__(self, rank: int, world: int, group=None):
        self.rank, self.world, self.group = int(rank), int(world), group

    # ---- collectives -----------------------------------------------------------------------------
    def all_reduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_max_(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    @staticmethod
    def _out(out, local, total_rows):
        shape = (total_rows,) + tuple(local.shape[1:])
        if out is not None and tuple(out.shape) == shape and out.dtype == local.dtype and out.is_contiguous():
            return out
        return torch.empty(shape, device=local.device, dtype=local.dtype)

    def all_gather_rows(self, local: torch.Tensor, total_rows: int, out=None) -> torch.Tensor:
        """Concatenate equally sized row blocks of all ranks (rank order) -> [total_rows, ...] (into `out` when it fits)."""
        if self.world == 1:
            return local
        out = self._out(out, local, total_rows)
        dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
        return out

    def all_gather_rows_async(self, local: torch.Tensor, total_rows: int, out=None):
        """Like all_gather_rows, but returns (out, work): the caller must `work.wait()` (a stream dependency, not a
        host block) before the first use of `out`; `work` is None when nothing is in flight."""
        if self.world == 1:
            return local, None
        out = self._out(out, local, total_rows)
        work = dist.all_gather_into_tensor(out, local.contiguous(), group=self.group, async_op=True)
        return out, work

    def reduce_scatter_rows(self, full: torch.Tensor) -> torch.Tensor:
        """Sum `full` [rows, ...] over the ranks and return this rank's equal row block [rows / world, ...]."""
        if self.world == 1:
            return full
        rows = full.shape[0]
        assert rows % self.world == 0
        out = torch.empty((rows // self.world,) + tuple(full.shape[1:]), device=full.device, dtype=full.dtype)
        dist.reduce_scatter_tensor(out, full.contiguous(), op=dist.ReduceOp.SUM, group=self.group)
        return out

    # ---- index arithmetic (pure host logic; covered by the gloo tests) ----------------------------------
    def row_range(self, n_total: int) -> Tuple[int, int]:
        """Contiguous global row range owned by this rank (the first n_total % world ranks get one more)."""
        base, rem = divmod(int(n_total), self.world)
        lo = self.rank * base + min(self.rank, rem)
        return lo, lo + base + (1 if self.rank < rem else 0)

    def local_counts(self, counts_per_component):
        """Global per-component sample counts -> (local counts per component, first global row)."""
        counts = [int(c) for c in counts_per_component]
        lo, hi = self.row_range(sum(counts))
        out, start = [], 0
        for c in counts:
            end = start + c
            out.append(max(0, min(hi, end) - max(lo, start)))
            start = end
        return out, lo

    def component_range(self, K: int) -> Optional[Tuple[int, int]]:
        """Equal component slice for the update step, or None when K is not divisible by world."""
        if K % self.world != 0:
            return None
        c = K // self.world
        return self.rank * c, (self.rank + 1) * c
