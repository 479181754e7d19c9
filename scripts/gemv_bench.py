"""BASELINE config 5: ExGEMV 32768 x 32768 fp64 on one B200 (CUDA events), 'N' and 'T'.
GB/s = (m*n + m + n) * 8 / t, GFLOP/s = 2*m*n / t, as the reference reports them (ExGEMV.cpp:208-211).

    python scripts/gemv_bench.py [m [kinds [trans]]]     kinds: comma list of naive,narrow,loguniform   trans: N,T
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exblas_b200 as xb
from exblas_b200 import common as cm
m = n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
kinds = (sys.argv[2] if len(sys.argv) > 2 else "naive,narrow,loguniform").split(",")
transs = (sys.argv[3] if len(sys.argv) > 3 else "N,T").split(",")
dev = torch.device("cuda:0")
h = xb.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream)
A = torch.empty(m * n, dtype=torch.float64, device=dev)
CH = 1 << 27


def fill(kind):
    for lo in range(0, m * n, CH):
        hi = min(m * n, lo + CH)
        if kind == "naive": A[lo:hi] = 1.1
        elif kind == "loguniform": A[lo:hi] = cm.init_fpuniform(m * n, 664, 332, seed=1, neg_ratio=2, lo=lo, hi=hi, device=dev)
        else: A[lo:hi] = cm.init_fpuniform(m * n, 10, 5, seed=1, neg_ratio=2, lo=lo, hi=hi, device=dev)     # the reference test's "10 0" range
    if kind == "naive":
        return torch.full((n,), 1.1, dtype=torch.float64, device=dev)
    return cm.init_fpuniform(n, 10, 5, seed=2, neg_ratio=2, device=dev)


def timed(trans, fpe, ee, K=5):
    y = torch.zeros(m, dtype=torch.float64, device=dev)
    for _ in range(2): xb.exgemv(trans, m, n, 1.0, A, m, 0, x, 1, 0, 0.0, y, 1, 0, fpe, ee, handle=h, sync=False)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(K): xb.exgemv(trans, m, n, 1.0, A, m, 0, x, 1, 0, 0.0, y, 1, 0, fpe, ee, handle=h, sync=False)
    e1.record(s); e1.synchronize()
    return e0.elapsed_time(e1) / K, y


for kind in kinds:
    x = fill(kind)
    torch.cuda.synchronize()
    for trans in transs:
        ref = None
        for label, fpe, ee, opts in [("fpe0 window", 0, False, {"window": 2}), ("fpe0 plain", 0, False, {"window": 0}),
                                     ("fpe3", 3, False, {"window": 2}), ("fpe8ee", 8, True, {"window": 2})] + \
                                    ([("fpe0 window 384 threads", 0, False, {"window": 2, "gemv_t_shape": 1})] if trans == "T" else
                                     [("fpe0 window 384 threads", 0, False, {"window": 2, "gemv_n_shape": 0})]):
            for k, v in opts.items(): h.set_option(k, v)
            ms, y = timed(trans, fpe, ee)
            same = True if ref is None else bool((y.view(torch.int64) == ref.view(torch.int64)).all())
            if ref is None: ref = y.clone()
            print(json.dumps({"op": "exgemv " + trans, "m": m, "n": n, "data": kind, "variant": label, "ms": round(ms, 3),
                              "GBs": round((m * n + m + n) * 8 / ms / 1e6, 1), "GFLOPs": round(2 * m * n / ms / 1e6, 1),
                              "bit_identical_to_first": same, "y0": float(y[0]), "status": h.last_status()}), flush=True)
        h.set_option("window", 2); h.set_option("gemv_t_shape", 2); h.set_option("gemv_n_shape", 1)
