"""ExGEMV 32768 x 32768 (BASELINE config 5): effect of the L2 bulk prefetch (TMA engine, option "gemv_prefetch" =
distance in rounds, 0 = off) on the window kernels, 'N' and 'T', narrow ("10 0") and naive data; alternating rounds."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exblas_b200 as xb
from exblas_b200 import common as cm
m = n = 32768
dev = torch.device("cuda:0")
h = xb.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream)
A = torch.empty(m * n, dtype=torch.float64, device=dev)
for lo in range(0, m * n, 1 << 27):
    A[lo:lo + (1 << 27)] = cm.init_fpuniform(m * n, 10, 5, seed=1, neg_ratio=2, lo=lo, hi=lo + (1 << 27), device=dev)
x = cm.init_fpuniform(n, 10, 5, seed=2, neg_ratio=2, device=dev)
y = torch.zeros(m, dtype=torch.float64, device=dev)

def timed(trans, K=8):
    for _ in range(2): xb.exgemv(trans, m, n, 1.0, A, m, 0, x, 1, 0, 0.0, y, 1, 0, 0, False, handle=h, sync=False)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(K): xb.exgemv(trans, m, n, 1.0, A, m, 0, x, 1, 0, 0.0, y, 1, 0, 0, False, handle=h, sync=False)
    e1.record(s); e1.synchronize()
    return e0.elapsed_time(e1) / K

ref = {}
for rnd in range(2):
    for trans in ("N", "T"):
        row = {"trans": trans, "round": rnd}
        for d in (0, 1, 2, 3, 4, 6, 8):
            h.set_option("gemv_prefetch", d)
            ms = timed(trans)
            key = y.clone()
            if trans not in ref: ref[trans] = key
            row[f"pf{d}"] = round((m * n + m + n) * 8 / ms / 1e6, 1)
            row["same"] = row.get("same", True) and bool((key.view(torch.int64) == ref[trans].view(torch.int64)).all())
        print(json.dumps(row), flush=True)
