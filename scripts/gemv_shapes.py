"""Development helper: launch-shape sweep of the ExGEMV window kernels (32768^2, narrow data)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exblas_b200 as xb
from exblas_b200 import common as cm
m = n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = torch.device("cuda:0")
h = xb.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream)
A = torch.empty(m * n, dtype=torch.float64, device=dev)
for lo in range(0, m * n, 1 << 27):
    hi = min(m * n, lo + (1 << 27))
    A[lo:hi] = cm.init_fpuniform(m * n, 10, 5, seed=1, neg_ratio=2, lo=lo, hi=hi, device=dev)
x = cm.init_fpuniform(n, 10, 5, seed=2, neg_ratio=2, device=dev)
y = torch.zeros(m, dtype=torch.float64, device=dev)
def timed(trans, K=8):
    for _ in range(2): xb.exgemv(trans, m, n, 1.0, A, m, 0, x, 1, 0, 0.0, y, 1, 0, 0, False, handle=h, sync=False)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(K): xb.exgemv(trans, m, n, 1.0, A, m, 0, x, 1, 0, 0.0, y, 1, 0, 0, False, handle=h, sync=False)
    e1.record(s); e1.synchronize()
    return e0.elapsed_time(e1) / K
for nshape in (0, 1, 2):
    h.set_option("gemv_n_shape", nshape)
    for parts in (0, 8, 16, 24):
        h.set_option("gemv_parts", parts)
        ms = timed("N")
        print(json.dumps({"trans": "N", "shape": nshape, "parts": parts, "ms": round(ms, 3), "GBs": round((m * n + m + n) * 8 / ms / 1e6, 1)}), flush=True)
h.set_option("gemv_parts", 0)
for tshape in (0, 1, 2, 3):
    h.set_option("gemv_t_shape", tshape)
    ms = timed("T")
    print(json.dumps({"trans": "T", "shape": tshape, "ms": round(ms, 3), "GBs": round((m * n + m + n) * 8 / ms / 1e6, 1)}), flush=True)
