"""Grid sharding across GPUs (one process per GPU) for the XC hot path.

The grid shards naturally by point ranges: stages 1-3 are per point, stage 4 and the VJP reduce
over the grid.  Each rank owns a fixed contiguous range of grid points; dm, theta
and the basis tables are replicated.  There is exactly one exchange per direction: an
all-reduce(sum) of the packed buffer ``[vmat (N*N) | excsum | nelec]`` after the forward and of
``[dm_bar (N*N) | theta_bar]`` after the reverse pass (NCCL over NVLink on GPUs; gloo in the CPU
tests).  The rank -> range map and the in-rank reduction order are fixed, so results are
bit-stable for a given world size.

Global ("NN-AmplitudeEncoding") functionals see the whole density vector and therefore do not
shard by grid; shard those by molecule (batch) instead.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

TILE = 128  # the kernels' grid-row tile: shard boundaries are multiples of it


def shard_range(ngrids: int, rank: int, world: int, tile: int = TILE):
    """Contiguous [lo, hi) of grid points owned by `rank`; boundaries are multiples of `tile`
    (except the global end) and the ranges partition [0, ngrids)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    ntiles = (ngrids + tile - 1) // tile
    base, rem = divmod(ntiles, world)
    t_lo = rank * base + min(rank, rem)
    t_hi = t_lo + base + (1 if rank < rem else 0)
    return min(ngrids, t_lo * tile), min(ngrids, t_hi * tile)


def shard_batch(nbatch: int, rank: int, world: int):
    """Round-robin molecule assignment for batched small-molecule workloads (config c4)."""
    return list(range(rank, nbatch, world))


def rank_world(group=None):
    """(rank, world) of the initialised process group, (0, 1) without one."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def all_reduce_packed(buf: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks of a packed output buffer; no-op without an initialised group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


class ShardedXC:
    """An XCContext over this rank's grid shard plus the two collectives of the path."""

    def __init__(self, ctx, group=None):
        self.ctx = ctx
        self.group = group

    def nr_rks_fwd(self, dm, theta, xctype="NN", hermi=0, **kw):
        if str(xctype) == "NN-AmplitudeEncoding":
            raise NotImplementedError("global functionals need the whole grid on one rank; shard by molecule")
        out, resid = self.ctx.nr_rks_fwd(dm, theta, xctype, hermi, **kw)
        all_reduce_packed(out, self.group)
        return out, resid

    def nr_rks_vjp(self, theta, resid, e_bar, v_bar, xctype="NN", hermi=0, **kw):
        bar = self.ctx.nr_rks_vjp(theta, resid, e_bar, v_bar, xctype, hermi, **kw)
        all_reduce_packed(bar, self.group)
        return bar
