"""CPU tests of the N>1 host logic with the gloo backend (world_size 2 and 3): sharding, the
exact limb all-reduce and the final rounding.  Per-rank limbs come from the oracle here (there is
no GPU); on the GPU box the same merge is exercised end to end by bench.py --gpus N and
tests/test_gpu_multi.py."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from exblas_b200 import common as cm, dist as xd
    from oracle.oracle import Oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        O = Oracle()
        out = []
        for kind in ("loguniform", "illcond", "cancel", "nan"):
            lo, hi = xd.shard_bounds(n, rank, world)
            if kind == "loguniform":
                a = cm.init_fpuniform(n, 664, 332, seed=3, neg_ratio=2, lo=lo, hi=hi)
            elif kind == "illcond":
                a = cm.init_ill_cond(n, 1e32, seed=3, lo=lo, hi=hi)
            elif kind == "cancel":
                a = cm.cancelling_pair(n, "sum")[lo:hi]
            else:
                a = cm.init_naive(n, lo=lo, hi=hi)
            status = 0
            if kind == "nan" and rank == world - 1:
                status = 1                                  # this rank met a NaN
            _, limbs = O.exsum(a, fpe=4 if rank % 2 else 0)  # ranks may even use different FPE sizes
            merged, st = xd.allreduce_limbs(limbs, status)
            v0 = xd.value_from(merged, st, 0)
            v1 = xd.value_from(merged, st, 1)
            out.append((kind, merged.tolist(), st, v0, v1))
        # a multi-rank reducer without a transport must refuse to reduce (it would return this rank's shard only)
        from exblas_b200._lib import ExblasB200Error
        red = xd.DistributedReducer(None)
        try:
            red.exsum_async(4, None)
            out.append(("no_transport", "did not raise"))
        except ExblasB200Error:
            out.append(("no_transport", "raised"))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_limb_allreduce(world, oracle):
    import torch.multiprocessing as mp
    from exblas_b200 import common as cm
    n = 6000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every rank has identical bits ...
    for r in range(1, world):
        assert str(results[r]) == str(results[0])
    # ... equal to the single-process result on the whole vector
    whole = {
        "loguniform": cm.init_fpuniform(n, 664, 332, seed=3, neg_ratio=2),
        "illcond": cm.init_ill_cond(n, 1e32, seed=3),
        "cancel": cm.cancelling_pair(n, "sum"),
    }
    for rec in results[0]:
        if rec[0] == "no_transport":
            assert rec[1] == "raised"
            continue
        kind, merged, st, v0, v1 = rec
        if kind == "nan":
            assert st == 1 and np.isnan(v0) and np.isnan(v1)
            continue
        w0, wl = oracle.exsum(whole[kind], fpe=0, round_mode=0)
        w1, _ = oracle.exsum(whole[kind], fpe=0, round_mode=1)
        assert merged == wl.tolist(), kind
        assert st == 0 and v0 == w0 and v1 == w1, kind
        if kind == "cancel":
            assert v1 == 1.5


def test_shard_bounds_cover_exactly():
    from exblas_b200.dist import shard_bounds
    for n in (0, 1, 7, 8, 1000, (1 << 30) + 3):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_bounds(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (l0, h0), (l1, h1) in zip(edges, edges[1:]):
                assert h0 == l1 and l0 <= h0
