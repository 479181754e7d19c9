"""Development helper: does a TMA-engine L2 prefetch of whole tiles help the ExSUM / ExDOT streaming kernel?
Device time per reduction (graph replay over 8 rotating buffers: never an L2 hit from a previous call) for
prefetch distances 0 (off), 1, 2, 4 tiles, sizes 2^22 .. 2^29, fpe 3 on log-uniform data."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exblas_b200 as xb
from exblas_b200 import common as cm
dev = torch.device("cuda:0")
h = xb.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream)
big = cm.init_fpuniform(1 << 27, 664, 332, seed=1, neg_ratio=1, device=dev)
big = torch.cat([big] * 8)        # 2^30 elements

def graph_us(n, fpe, nbuf, K=8):
    nbuf = max(1, min(nbuf, big.numel() // n))
    views = [big[i * n:(i + 1) * n] for i in range(nbuf)]
    for v in views: h.exsum_async(n, v, 1, 0, fpe, False)
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for k in range(K): h.exsum_async(n, views[k % nbuf], 1, 0, fpe, False)
    g.replay(); s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(5): g.replay()
    e1.record(s); e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * K)

for rnd in range(2):
    for lg in (22, 24, 25, 26, 28, 30):
        n = 1 << lg
        row = {"log2n": lg, "round": rnd}
        for d in (0, 1, 2, 4):
            h.set_option("reduce_prefetch", d)
            us = graph_us(n, 3, 8)
            row[f"pf{d}"] = {"us": round(us, 2), "GBs": round(n * 8 / us / 1e3, 1)}
        print(json.dumps(row), flush=True)
