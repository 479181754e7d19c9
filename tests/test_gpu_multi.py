"""Multi-GPU parity (pytest -m gpu on a box with >= 2 GPUs; skipped otherwise): the sharded reduction
with the C-ABI NCCL limb all-reduce returns, on every rank, the bits of the single-GPU result."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
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
    import torch
    import torch.distributed as dist
    import exblas_b200 as xb
    from exblas_b200 import common as cm, dist as xd
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        h = xb.Handle(rank)
        red = xd.DistributedReducer(h)
        red.init_nccl()
        lo, hi = xd.shard_bounds(n, rank, world)
        out = []
        a = cm.init_fpuniform(n, 664, 332, seed=9, neg_ratio=2, lo=lo, hi=hi, device=f"cuda:{rank}")
        x, y = cm.cancelling_pair(n, "dot", device=f"cuda:{rank}")
        x, y = x[lo:hi].contiguous(), y[lo:hi].contiguous()
        torch.cuda.synchronize()
        for fpe, ee in [(0, False), (4, False), (8, True)]:
            for rm in (0, 1):
                v, l, st = red.exsum(hi - lo, a, fpe, ee, rm)
                out.append(("sum", fpe, ee, rm, v, l.tolist(), st))
            v, l, st = red.exdot(hi - lo, x, y, fpe, ee, 1)
            out.append(("dot", fpe, ee, 1, v, l.tolist(), st))
        # the FUSED transport (limb exchange inside the closing kernel over peer memory) must give the
        # same bits as the NCCL transport, reduction after reduction (epochs alternate mailbox sets)
        nccl_results = [(rec[4], rec[5], rec[6]) for rec in out]
        red.init_peer()
        fused = []
        for rep in range(3):
            for fpe, ee in [(0, False), (4, False), (8, True)]:
                for rm in (0, 1):
                    v, l, st = red.exsum(hi - lo, a, fpe, ee, rm)
                    fused.append((v, l.tolist(), st))
                v, l, st = red.exdot(hi - lo, x, y, fpe, ee, 1)
                fused.append((v, l.tolist(), st))
        assert fused == nccl_results * 3, "fused peer-memory exchange differs from the NCCL all-reduce"
        out.append(("fused_ok", len(fused)))
        # a rank-local special must reach every rank (fused transport, then NCCL again)
        if rank == world - 1:
            a[5] = float("inf")
        v, l, st = red.exsum(hi - lo, a, 4, False, 0)
        out.append(("inf", v, st))
        h.set_option("fused_allreduce", 0)
        v, l, st = red.exsum(hi - lo, a, 4, False, 0)
        out.append(("inf", v, st))
        q.put((rank, out))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_equals_single_gpu(gpu):
    import torch
    import torch.multiprocessing as mp
    from exblas_b200 import common as cm
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 10_000_002
    for world in sorted({2, min(ngpu, 4), min(ngpu, 8)}):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
        for p in procs:
            p.start()
        results = {}
        import queue as _q
        import time as _t
        t0 = _t.time()
        while len(results) < world:
            try:
                r, out = q.get(timeout=5)
                results[r] = out
            except _q.Empty:
                dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
                assert not dead, f"a rank died with exit code {dead}"
                assert _t.time() - t0 < 600, "timeout waiting for ranks"
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
        for r in range(1, world):
            assert str(results[r]) == str(results[0]), f"rank {r} differs from rank 0 (world {world})"
        whole = cm.init_fpuniform(n, 664, 332, seed=9, neg_ratio=2, device="cuda")
        for rec in results[0]:
            if rec[0] == "sum":
                _, fpe, ee, rm, v, l, st = rec
                gpu.exsum_async(n, whole, 1, 0, fpe, ee, rm)
                w, wl, wst = gpu.fetch()
                assert v == w and l == wl.tolist() and st == wst == 0, (world, fpe, ee, rm)
            elif rec[0] == "dot":
                assert rec[4] == 1.5 and rec[6] == 0, (world, rec[:4])
            elif rec[0] == "fused_ok":
                assert rec[1] == 27
            else:
                assert rec[1] == float("inf") and rec[2] == 2
