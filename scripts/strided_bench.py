"""Strided ExSUM / ExDOT (inca != 1) on one B200: GB/s of USEFUL data (8 B per summand)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exblas_b200 as xb
from exblas_b200 import common as cm
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 26)
dev = torch.device("cuda:0")
h = xb.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream)
for inc in (1, 2, 3, 4, 8):
    a = cm.init_fpuniform(n * inc, 664, 332, seed=1, neg_ratio=2, device=dev)
    b = cm.init_fpuniform(n * inc, 10, 5, seed=2, neg_ratio=2, device=dev)
    for op, fpe in (("sum", 0), ("sum", 3), ("dot", 0), ("dot", 3)):
        def run():
            if op == "sum": h.exsum_async(n, a, inc, 0, fpe, False, 0)
            else: h.exdot_async(n, a, inc, 0, b, inc, 0, fpe, False, 0)
        for _ in range(2): run()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5): run()
        e1.record(s); e1.synchronize()
        ms = e0.elapsed_time(e1) / 5
        v, _, st = h.fetch()
        print(json.dumps({"op": "ex" + op, "n": n, "inc": inc, "fpe": fpe, "ms": round(ms, 3),
                          "useful_GBs": round(n * (8 if op == "sum" else 16) / ms / 1e6, 1),
                          "touched_GBs": round(n * inc * (8 if op == "sum" else 16) / ms / 1e6, 1), "value": v, "status": st}), flush=True)
