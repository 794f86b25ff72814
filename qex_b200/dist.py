"""Grid sharding across GPUs (one process per GPU) for the XC hot path.

The grid shards naturally by point ranges: stages 1-3 are per point, stage 4 and the VJP reduce
over the grid.  Each rank owns a fixed contiguous range of grid points; dm, theta
and the basis tables are replicated.  There is exactly one exchange per direction: an
all-reduce(sum) of the packed buffer ``[vmat (N*N) | excsum | nelec]`` after the forward and of
``[dm_bar (N*N) | theta_bar]`` after the reverse pass.  On GPUs the collectives go through the C ABI
(``qexxc_allreduce`` / ``qexxc_bcast`` on an NCCL communicator, NVLink/NVSwitch); the CPU tests of the
host logic use a ``torch.distributed`` gloo group instead.  The rank -> range map and the in-rank
reduction order are fixed, so results are bit-stable for a given world size.

Global ("NN-AmplitudeEncoding") functionals see the whole density vector and therefore do not
shard by grid; shard those by molecule (batch) instead.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
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


class Comm:
    """An NCCL communicator behind the C ABI (``qexxc_comm_create`` / ``qexxc_allreduce`` / ``qexxc_bcast``).

    The 128-byte NCCL unique id is produced by rank 0 and shipped to the other ranks over the already
    initialised ``torch.distributed`` group (any backend; it is only the side channel).  ``world == 1``
    needs no NCCL and makes every collective a no-op."""

    def __init__(self, rank: int | None = None, world: int | None = None, device: int | None = None, group=None):
        from . import _lib

        self._lib = _lib
        self.lib = _lib.load()
        r, w = rank_world(group)
        self.rank = r if rank is None else int(rank)
        self.world = w if world is None else int(world)
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._h = C.c_void_p()
        if self.world > 1:
            uid = (C.c_ubyte * 128)()
            if self.rank == 0:
                _lib.check(self.lib.qexxc_comm_unique_id(uid))
            box = [bytes(uid)]
            dist.broadcast_object_list(box, src=0, group=group)
            uid = (C.c_ubyte * 128).from_buffer_copy(box[0])
            _lib.check(self.lib.qexxc_comm_create(C.byref(self._h), self.device, self.world, self.rank, uid))

    @staticmethod
    def _st(stream):
        s = torch.cuda.current_stream() if stream is None else stream
        return C.c_void_p(s.cuda_stream)

    def all_reduce(self, t: torch.Tensor, stream=None) -> torch.Tensor:
        if self.world > 1:
            assert t.dtype == torch.float64 and t.is_contiguous() and t.is_cuda
            self._lib.check(self.lib.qexxc_allreduce(self._h, C.c_void_p(t.data_ptr()), t.numel(), self._st(stream)))
        return t

    def bcast(self, t: torch.Tensor, root: int = 0, stream=None) -> torch.Tensor:
        if self.world > 1:
            assert t.dtype == torch.float64 and t.is_contiguous() and t.is_cuda
            self._lib.check(self.lib.qexxc_bcast(self._h, C.c_void_p(t.data_ptr()), t.numel(), int(root),
                                                 self._st(stream)))
        return t

    @property
    def calls(self) -> int:
        return int(self.lib.qexxc_comm_calls(self._h)) if self._h.value else 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.qexxc_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedXC:
    """An XCContext over this rank's grid shard plus the collectives of the path.

    ``comm`` (a :class:`Comm`) routes the collectives through the C ABI; without it they go through
    ``torch.distributed`` on ``group`` (the gloo CPU tests, or a caller that already owns a process group)."""

    def __init__(self, ctx, group=None, comm: Comm | None = None):
        self.ctx = ctx
        self.group = group
        self.comm = comm
        self.rank, self.world = (comm.rank, comm.world) if comm is not None else rank_world(group)
        self._side = self._s_in = None
        self.timing = None  # optional dict of CUDA events of the last step()

    # ---- collectives -----------------------------------------------------------------------
    def _all_reduce(self, t, stream=None):
        if self.comm is not None:
            return self.comm.all_reduce(t, stream)
        if stream is not None and stream != torch.cuda.current_stream():
            with torch.cuda.stream(stream):
                return all_reduce_packed(t, self.group)
        return all_reduce_packed(t, self.group)

    def _bcast(self, t, stream=None):
        if self.world == 1:
            return t
        if self.comm is not None:
            return self.comm.bcast(t, 0, stream)
        if stream is not None and stream != torch.cuda.current_stream():
            with torch.cuda.stream(stream):
                dist.broadcast(t, src=0, group=self.group)
        else:
            dist.broadcast(t, src=0, group=self.group)
        return t

    def _streams(self):
        if self._side is None:
            dev = self.ctx.tdev
            self._side = torch.cuda.Stream(device=dev, priority=-1)  # collectives / D2H of the forward result
            self._s_in = torch.cuda.Stream(device=dev, priority=-1)  # H2D + broadcast of the cotangents
        return self._side, self._s_in

    # ---- the two calls of the path, each followed by its collective ------------------------
    def nr_rks_fwd(self, dm, theta, xctype="NN", hermi=0, **kw):
        if str(xctype) == "NN-AmplitudeEncoding":
            raise NotImplementedError("global functionals need the whole grid on one rank; shard by molecule")
        out, resid = self.ctx.nr_rks_fwd(dm, theta, xctype, hermi, **kw)
        self._all_reduce(out)
        return out, resid

    def nr_rks_vjp(self, theta, resid, e_bar, v_bar, xctype="NN", hermi=0, **kw):
        bar = self.ctx.nr_rks_vjp(theta, resid, e_bar, v_bar, xctype, hermi, **kw)
        self._all_reduce(bar)
        return bar

    # ---- fwd + VJP as one pipelined step ----------------------------------------------------
    def step(self, dm, theta, e_bar, v_bar, xctype, out, bar, resid, hermi=0, record=False):
        """nr_rks forward + its VJP on device-resident inputs.  The forward all-reduce runs on a side stream
        and overlaps the first kernels of the VJP (pad_sym + rowquad do not read `out`); the VJP all-reduce
        closes the step on the current stream.  Returns (out, bar), both summed over ranks."""
        ctx = self.ctx
        main = torch.cuda.current_stream()
        ctx.nr_rks_fwd(dm, theta, xctype, hermi, out=out, resid=resid)
        ev = None
        if self.world > 1:
            side, _ = self._streams()
            side.wait_stream(main)
            if record:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                ev[0].record(side)
            self._all_reduce(out, side)
            if record:
                ev[1].record(side)
        ctx.nr_rks_vjp(theta, resid, e_bar, v_bar, xctype, hermi, out=bar)
        if self.world > 1:
            if record:
                ev[2].record(main)
            self._all_reduce(bar, main)
            if record:
                ev[3].record(main)
                self.timing = ev
            main.wait_stream(side)
        return out, bar

    def collective_ms(self):
        """(forward all-reduce ms, VJP all-reduce ms) of the last ``step(record=True)``; synchronises."""
        if not self.timing:
            return 0.0, 0.0
        ev = self.timing
        ev[3].synchronize()
        ev[1].synchronize()
        return ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])

    # ---- host buffers in, host buffers out ---------------------------------------------------
    def make_host_io(self, nbatch, n_theta):
        """Pinned host staging + device mirrors for step_host: inp = [dm | theta | e_bar | v_bar]."""
        N, B = self.ctx.nao, nbatch
        n1 = B * N * N + n_theta
        n2 = B + B * N * N
        io = {
            "n1": n1, "n2": n2, "B": B, "n_theta": n_theta,
            "d_inp": self.ctx.empty(n1 + n2),
            "out": self.ctx.empty(B, N * N + 2), "bar": self.ctx.empty(B * N * N + n_theta),
            "resid": self.ctx.empty(self.ctx.resid_doubles),
        }
        if self.rank == 0:
            io["h_inp"] = torch.empty(n1 + n2, dtype=torch.float64).pin_memory()
            io["h_out"] = torch.empty(B, N * N + 2, dtype=torch.float64).pin_memory()
            io["h_bar"] = torch.empty(B * N * N + n_theta, dtype=torch.float64).pin_memory()
        return io

    def step_host(self, io, h_coords, h_weights, d_coords, d_weights, xctype, deriv=0, hermi=0):
        """One whole call of the drop-in with HOST buffers: rank 0 holds (dm, theta, e_bar, v_bar) in
        io["h_inp"] (pinned); every rank holds its own grid shard (h_coords, h_weights, pinned).
        Rank 0 uploads the replicated inputs once and they reach the peers by NCCL broadcast over NVLink; the
        cotangents travel on a side stream while the forward computes; the forward result is reduced and
        downloaded (rank 0 only) while the VJP computes.  Synchronises at the end; results in io["h_out"],
        io["h_bar"] on rank 0."""
        ctx, N, B = self.ctx, self.ctx.nao, io["B"]
        n1, n2, nth = io["n1"], io["n2"], io["n_theta"]
        main = torch.cuda.current_stream()
        side, s_in = self._streams()
        d_inp = io["d_inp"]
        s_in.wait_stream(main)
        side.wait_stream(main)
        # grid shard of this rank first (the AO evaluation waits for it; H2D copies share one engine)
        d_coords.copy_(h_coords, non_blocking=True)
        d_weights.copy_(h_weights, non_blocking=True)
        # the replicated (dm | theta), then the cotangents (needed by the VJP only), travel on a side stream: upload on rank
        # 0, NCCL broadcast over NVLink -- under this rank's grid upload, AO evaluation and digit slicing, which need neither
        with torch.cuda.stream(s_in):
            if self.rank == 0:
                d_inp[:n1].copy_(io["h_inp"][:n1], non_blocking=True)
            self._bcast(d_inp[:n1], s_in)
            ev_dm = torch.cuda.Event()
            ev_dm.record(s_in)
            if self.rank == 0:
                d_inp[n1:].copy_(io["h_inp"][n1:], non_blocking=True)
            self._bcast(d_inp[n1:], s_in)
        dm = d_inp[: B * N * N].view(B, N, N)
        theta = d_inp[B * N * N : n1]
        e_bar = d_inp[n1 : n1 + B]
        v_bar = d_inp[n1 + B :].view(B, N, N)
        ctx.set_grid(d_coords, d_weights)
        ctx.eval_ao(deriv)
        ctx.prepare_contractions()
        main.wait_event(ev_dm)
        ctx.nr_rks_fwd(dm, theta, xctype, hermi, out=io["out"], resid=io["resid"])
        side.wait_stream(main)
        side.wait_stream(s_in)  # keeps NCCL's per-communicator issue order identical to the stream order
        self._all_reduce(io["out"], side)
        if self.rank == 0:
            with torch.cuda.stream(side):
                io["h_out"].copy_(io["out"], non_blocking=True)
        main.wait_stream(s_in)
        ctx.nr_rks_vjp(theta, io["resid"], e_bar, v_bar, xctype, hermi, out=io["bar"])
        main.wait_stream(side)
        self._all_reduce(io["bar"], main)
        if self.rank == 0:
            io["h_bar"].copy_(io["bar"], non_blocking=True)
        torch.cuda.synchronize()
        return (io["h_out"], io["h_bar"]) if self.rank == 0 else (None, None)

    @staticmethod
    def pack_host_inputs(io, dm, theta, e_bar, v_bar):
        """Fill io["h_inp"] (rank 0) from numpy arrays."""
        B, n1 = io["B"], io["n1"]
        h = io["h_inp"].numpy()
        nn = np.asarray(dm).size
        h[:nn] = np.asarray(dm, dtype=np.float64).ravel()
        h[nn:n1] = np.asarray(theta, dtype=np.float64).ravel()
        h[n1 : n1 + B] = np.broadcast_to(np.asarray(e_bar, dtype=np.float64).ravel(), (B,))
        h[n1 + B :] = np.broadcast_to(np.asarray(v_bar, dtype=np.float64), (B,) + np.asarray(dm).shape[-2:]).ravel()
