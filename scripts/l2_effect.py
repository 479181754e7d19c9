"""Development helper: is the 2^24 anomaly an L2 effect?  Time reductions of 2^22..2^26 elements (graph replay) over
ONE buffer (replays may hit in the 126 MB L2) and rotating over 8 distinct buffers (never)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exblas_b200 as xb
from exblas_b200 import common as cm
dev = torch.device("cuda:0")
h = xb.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream)
big = cm.init_fpuniform(1 << 27, 664, 332, seed=1, neg_ratio=1, device=dev)
big = torch.cat([big] * 4)        # 2^29 elements = 4 GiB

def graph_us(n, fpe, nbuf, K=16):
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

for lg in (20, 22, 23, 24, 25, 26):
    n = 1 << lg
    row = {"log2n": lg}
    for fpe in (0, 3):
        one = graph_us(n, fpe, 1); rot = graph_us(n, fpe, 8)
        row[f"fpe{fpe}"] = {"one_buffer_us": round(one, 2), "rotating_us": round(rot, 2), "rot_GBs": round(n * 8 / rot / 1e3, 1)}
    print(json.dumps(row), flush=True)
