"""Small-n latency with launch overhead amortised: capture 50 exsum_async calls into one CUDA graph
(the device entry point is capture-safe: one kernel launch, no allocation, no synchronisation) and
time replays.  Prints device microseconds per reduction."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exblas_b200 as xb
from exblas_b200 import common as cm
from oracle.oracle import Oracle
dev = torch.device("cuda:0")
h = xb.Handle(0)
a = cm.init_fpuniform(1 << 22, 664, 332, seed=1, neg_ratio=2, device=dev)
host = a.cpu().numpy()
O = Oracle()
s = torch.cuda.Stream()
h.set_stream(s.cuda_stream)
K = 50
for lg in (10, 12, 14, 16, 18, 20, 22):
    n = 1 << lg
    row = {"log2n": lg}
    for fpe, ee, tag in ((0, False, "fpe0"), (8, True, "fpe8ee")):
        with torch.cuda.stream(s):
            for _ in range(3): h.exsum_async(n, a, 1, 0, fpe, ee)      # warm-up outside capture (sets function attributes)
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(K): h.exsum_async(n, a, 1, 0, fpe, ee)
            g.replay(); s.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(10): g.replay()
            e1.record(s); e1.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (10 * K)
        v, limbs, st = h.fetch()
        ok = v == O.exsum(host[:n], fpe=0)[0]
        row[tag] = {"graph_us_per_call": round(us, 2), "GBs": round(n * 8 / us / 1e3, 1), "bit_exact": bool(ok)}
    print(json.dumps(row), flush=True)
