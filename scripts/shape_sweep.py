"""Development helper: device time per reduction (CUDA-graph replay of 20 captured calls) for hand-set launch shapes,
to place the size thresholds of the automatic shape (solo_max_elems / small_max_elems).
    python scripts/shape_sweep.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exblas_b200 as xb
from exblas_b200 import common as cm
dev = torch.device("cuda:0")
h = xb.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream)
a = cm.init_fpuniform(1 << 26, 664, 332, seed=1, neg_ratio=1, device=dev)

def graph_us(n, fpe, ee, K=20):
    for _ in range(3): h.exsum_async(n, a, 1, 0, fpe, ee)
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(K): h.exsum_async(n, a, 1, 0, fpe, ee)
    g.replay(); s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(5): g.replay()
    e1.record(s); e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * K)

for lg in (12, 13, 14, 15, 16, 18, 20, 22, 23, 24, 26):
    n = 1 << lg
    row = {"log2n": lg}
    for fpe, ee in ((0, False), (3, False)):
        r = {}
        h.set_option("auto_shape", 1)
        h.set_option("solo_max_elems", 1 << 13)
        r["auto"] = round(graph_us(n, fpe, ee), 2)
        if lg <= 16:
            h.set_option("solo_max_elems", 1 << 16)
            r["solo"] = round(graph_us(n, fpe, ee), 2)
            h.set_option("solo_max_elems", 1 << 13)
        for T in (128, 256, 512):
            h.set_option("block_threads", T)
            h.set_option("blocks", 0)
            r[f"T{T}"] = round(graph_us(n, fpe, ee), 2)
        h.set_option("auto_shape", 1)
        row[f"fpe{fpe}"] = r
    print(json.dumps(row), flush=True)
