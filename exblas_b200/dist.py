"""Multi-GPU ExSUM / ExDOT: one process per GPU, contiguous shards, exact limb all-reduce.

B200 restatement of the reference's MPI path (src/cpu/blas/blas1/ExSUM.cpp:33-65 scatter,
:266-273 `MPI_Reduce(MPI_LONG, MPI_SUM)` of normalised limbs, then `Round()`): every rank reduces
its shard to 39 normalised int64 limbs on its GPU, the ranks sum those limbs as integers (exact,
order-free), and every rank normalises and rounds the same integers -- so all ranks, and any
number of ranks, return identical bits.

Three transports for the message (39 limbs + status flags; 44 x int64 for the all-reduce variants):
  * `init_peer()`            -- FUSED: the closing reduction kernel stores the message into every
    peer's mailbox over NVLink peer memory and merges what it receives, no collective call at all
    (exblas_b200_peer_export / exblas_b200_peer_attach);
  * `Handle.allreduce_async` -- ncclAllReduce on the handle's own stream through the C ABI
    (exblas_b200_comm_init / exblas_b200_allreduce_async); used by bench.py and on GPUs;
  * `allreduce_limbs`        -- torch.distributed.all_reduce on a tensor, any backend (gloo in
    the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import blas1
from ._lib import LIMBS, ROUND_REFERENCE

FLAG_SLOTS = 5


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous element range of `rank`: [g*ceil(n/G), min(n, (g+1)*ceil(n/G)))."""
    per = -(-n // world)
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


def pack(limbs, status: int) -> np.ndarray:
    """limbs[39] + one counter per status flag, so that flags survive an integer sum."""
    msg = np.zeros(LIMBS + FLAG_SLOTS, dtype=np.int64)
    msg[:LIMBS] = np.asarray(limbs, dtype=np.int64)
    for k in range(FLAG_SLOTS):
        msg[LIMBS + k] = (status >> k) & 1
    return msg


def unpack(msg) -> Tuple[np.ndarray, int]:
    msg = np.asarray(msg, dtype=np.int64)
    status = 0
    for k in range(FLAG_SLOTS):
        if msg[LIMBS + k] != 0:
            status |= 1 << k
    limbs, _ = blas1.normalize_limbs(msg[:LIMBS])
    return limbs, status


def allreduce_limbs(limbs, status: int = 0, group=None) -> Tuple[np.ndarray, int]:
    """Exact combination of per-rank limbs over torch.distributed (any backend).  Each rank's limbs
    must be normalised (limbs 0..37 < 2^52), so the sum over <= 2048 ranks cannot overflow."""
    import torch
    import torch.distributed as dist
    msg = torch.from_numpy(pack(limbs, status))
    if dist.get_backend(group) == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        t = msg.to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        msg = t.cpu()
    else:
        dist.all_reduce(msg, op=dist.ReduceOp.SUM, group=group)
    return unpack(msg.numpy())


def value_from(limbs, status: int, round_mode: int = ROUND_REFERENCE) -> float:
    """IEEE result for the specials, else the chosen rounding of the exact sum."""
    from ._lib import ST_NAN, ST_NEGINF, ST_POSINF
    if (status & ST_NAN) or ((status & ST_POSINF) and (status & ST_NEGINF)):
        return float("nan")
    if status & ST_POSINF:
        return float("inf")
    if status & ST_NEGINF:
        return float("-inf")
    return blas1.round_limbs(limbs, round_mode)


class DistributedReducer:
    """Rank-local handle + communicator.  `init_nccl()` must be called collectively once."""

    def __init__(self, handle: Optional[blas1.Handle] = None, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.handle = handle
        self._nccl_ready = False
        self.fused = False
        if handle is not None and self.world > 1:
            handle.set_option("world_size", self.world)      # the C ABI then refuses to finish a reduction locally

    def init_nccl(self) -> None:
        """Create the C-ABI NCCL communicator: rank 0 makes the unique id, everyone gets it through
        torch.distributed, then all ranks call ncclCommInitRank (collective)."""
        if self.world == 1 or self._nccl_ready:
            return
        obj = [blas1.nccl_unique_id() if self.rank == 0 else None]
        self.dist.broadcast_object_list(obj, src=0, group=self.group)
        self.handle.comm_init(self.world, self.rank, obj[0])
        self._nccl_ready = True

    def init_peer(self) -> bool:
        """Set up the FUSED exchange: every rank exports its mailbox (CUDA IPC), the handles are gathered
        in rank order, every rank maps its peers.  Afterwards each reduction exchanges its limbs inside
        the closing kernel over NVLink peer memory; no NCCL call, no extra launch.  Collective-safe:
        returns False on EVERY rank (and leaves the handle on its previous transport) if any rank could
        not export or map a mailbox, e.g. where CUDA IPC is not permitted."""
        if self.world == 1:
            return True
        try:
            mine = self.handle.peer_export()
        except Exception:
            mine = None
        handles = [None] * self.world
        self.dist.all_gather_object(handles, mine, group=self.group)
        ok = all(hd is not None for hd in handles)
        if ok:
            try:
                self.handle.peer_attach(self.world, self.rank, handles)
            except Exception:
                ok = False
        flags = [None] * self.world
        self.dist.all_gather_object(flags, ok, group=self.group)   # also: nobody launches before all are mapped
        ok = all(flags)
        self.handle.set_option("fused_allreduce", 1 if ok else 0)
        self.fused = ok
        if not ok:
            self.init_nccl()             # never leave a multi-rank reducer without a transport (collective: all ranks get here)
        return ok

    def _require_transport(self) -> None:
        if self.world > 1 and not (self.fused or self._nccl_ready):
            raise blas1.ExblasB200Error(
                f"DistributedReducer: {self.world} ranks but no transport -- call init_peer() or init_nccl() "
                "collectively first (a reduction now would return this rank's shard only)")

    # device-resident shard in, identical value on every rank out
    def exsum(self, n_local: int, d_a, fpe: int = 0, early_exit: bool = False, round_mode: int = ROUND_REFERENCE):
        self.exsum_async(n_local, d_a, fpe, early_exit, round_mode)
        return self.handle.fetch()

    def exsum_async(self, n_local, d_a, fpe=0, early_exit=False, round_mode=ROUND_REFERENCE):
        self._require_transport()
        self.handle.exsum_async(n_local, d_a, 1, 0, fpe, early_exit, round_mode)
        if self.world > 1:
            self.handle.allreduce_async(round_mode)

    def exdot(self, n_local, d_a, d_b, fpe=0, early_exit=False, round_mode=ROUND_REFERENCE):
        self.exdot_async(n_local, d_a, d_b, fpe, early_exit, round_mode)
        return self.handle.fetch()

    def exdot_async(self, n_local, d_a, d_b, fpe=0, early_exit=False, round_mode=ROUND_REFERENCE):
        self._require_transport()
        self.handle.exdot_async(n_local, d_a, 1, 0, d_b, 1, 0, fpe, early_exit, round_mode)
        if self.world > 1:
            self.handle.allreduce_async(round_mode)
