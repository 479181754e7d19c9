"""BASELINE config 5: ExGEMV 32768 x 32768 fp64 non-transpose on one B200 (CUDA events).
GB/s = (m*n + m + n) * 8 / t, GFLOP/s = 2*m*n / t, as the reference reports them (ExGEMV.cpp:208-211)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exblas_b200 as xb
from exblas_b200 import common as cm
m = n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
kind = sys.argv[2] if len(sys.argv) > 2 else "naive"
dev = torch.device("cuda:0")
h = xb.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream)
A = torch.empty(m * n, dtype=torch.float64, device=dev)
CH = 1 << 27
for lo in range(0, m * n, CH):
    hi = min(m * n, lo + CH)
    if kind == "naive": A[lo:hi] = 1.1
    elif kind == "loguniform": A[lo:hi] = cm.init_fpuniform(m * n, 664, 332, seed=1, neg_ratio=2, lo=lo, hi=hi, device=dev)
    else: A[lo:hi] = cm.init_fpuniform(m * n, 40, 20, seed=1, neg_ratio=2, lo=lo, hi=hi, device=dev)
x = cm.init_fpuniform(n, 40, 20, seed=2, neg_ratio=2, device=dev) if kind != "naive" else torch.full((n,), 1.1, dtype=torch.float64, device=dev)
y = torch.zeros(m, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
for fpe, ee in [(0, False), (3, False), (4, False), (8, False), (4, True), (8, True)]:
    for parts in ([0] if fpe else [0, 5, 9, 18, 37]):
        h.set_option("gemv_parts", parts)
        for _ in range(2): xb.exgemv("N", m, n, 1.0, A, m, 0, x, 1, 0, 0.0, y, 1, 0, fpe, ee, handle=h, sync=False)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s)
        K = 5
        for _ in range(K): xb.exgemv("N", m, n, 1.0, A, m, 0, x, 1, 0, 0.0, y, 1, 0, fpe, ee, handle=h, sync=False)
        e1.record(s); e1.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(json.dumps({"op": "exgemv N", "m": m, "n": n, "data": kind, "fpe": fpe, "early_exit": ee, "parts": parts, "ms": round(ms, 3),
                          "GBs": round((m * n + m + n) * 8 / ms / 1e6, 1), "GFLOPs": round(2 * m * n / ms / 1e6, 1), "y0": float(y[0])}), flush=True)
