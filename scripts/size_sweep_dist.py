"""BASELINE config 4 at N GPUs: ExSUM size sweep, TOTAL size 2^10 .. 2^32 doubles sharded over the ranks
(strong scaling: the latency-bound to HBM-bound crossover), fused peer-memory limb exchange inside the kernel.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/size_sweep_dist.py [max_log2n]

One JSON line per total size (rank 0): device time per reduction (max over ranks, CUDA events around back-to-back
collective reductions) and aggregate GB/s; every result is checked to be identical on all ranks."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import exblas_b200 as xb
from exblas_b200 import common as cm, dist as xd

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
h = xb.Handle(lr)
s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream)
transport = "single"
red = None
if world > 1:
    red = xd.DistributedReducer(h)
    red.init_nccl()
    transport = "fused peer-memory exchange" if red.init_peer() else "nccl"
max_lg = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nmax_local = (1 << max_lg) // world
a = torch.empty(nmax_local, dtype=torch.float64, device=dev)
CH = 1 << 27
for lo in range(0, nmax_local, CH):
    hi = min(nmax_local, lo + CH)
    a[lo:hi] = cm.init_fpuniform(1 << max_lg, 664, 332, seed=1, neg_ratio=2, lo=rank * nmax_local + lo, hi=rank * nmax_local + hi, device=dev)
torch.cuda.synchronize()
for lg in range(10, max_lg + 1, 2):
    n_local = max((1 << lg) // world, 1)
    row = {"log2n_total": lg, "n_gpus": world, "transport": transport}
    for fpe, ee, tag in [(0, False, "fpe0"), (3, False, "fpe3"), (8, True, "fpe8ee")]:
        reps = 200 if lg <= 20 else (20 if lg <= 26 else 5)
        def one():
            h.exsum_async(n_local, a, 1, 0, fpe, ee)
            if world > 1: h.allreduce_async(0)
        for _ in range(3): one()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps): one()
        e1.record(s); e1.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        v, limbs, st = h.fetch()
        same = True
        if world > 1:
            t = torch.tensor([us], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            us = float(t.item())
            lt = torch.tensor(limbs.tolist(), dtype=torch.int64, device=dev)
            l0 = lt.clone(); dist.broadcast(l0, src=0)
            ok = torch.tensor([int(bool((lt == l0).all()))], device=dev); dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            same = bool(ok.item())
        row[tag] = {"device_us": round(us, 2), "GBs": round(n_local * world * 8 / us / 1e3, 1), "identical_on_all_ranks": same}
    if rank == 0:
        print(json.dumps(row), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
