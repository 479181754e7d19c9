"""Development helper: in-process sweep over handle options (prefetch distance, adaptive, block size)."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exblas_b200 as xb
from exblas_b200 import common as cm
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = 1 << lg
dev = torch.device("cuda:0")
h = xb.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream)
data = {"logu": cm.init_fpuniform(n, 664, 332, seed=1, neg_ratio=2, device=dev), "naive": cm.init_naive(n, device=dev)}
data["ill"] = cm.init_ill_cond(n, 1e32, seed=1, device=dev)
data["logupos"] = cm.init_fpuniform(n, 664, 332, seed=1, device=dev)
b = cm.init_ill_cond(n, 1e32, seed=2, device=dev)
torch.cuda.synchronize()
def timeit(fn, reps=5):
    fn(); fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps): fn()
    e1.record(s); e1.synchronize()
    return e0.elapsed_time(e1) / reps
cfgs = [("exsum", "logupos", 0, 0), ("exsum", "logupos", 3, 0), ("exsum", "logupos", 8, 0), ("exsum", "logu", 0, 0), ("exsum", "logu", 3, 0), ("exsum", "logu", 4, 0), ("exsum", "logu", 8, 0), ("exsum", "naive", 2, 0), ("exsum", "naive", 3, 0), ("exsum", "naive", 4, 0),
        ("exsum", "naive", 8, 0), ("exsum", "naive", 8, 1), ("exsum", "ill", 8, 1), ("exdot", "ill", 0, 0), ("exdot", "ill", 3, 0), ("exdot", "ill", 8, 0), ("exdot", "ill", 8, 1)]
for pf in [1]:
    h.set_option("adaptive", pf)
    row = []
    for op, kind, fpe, ee in cfgs:
        a = data[kind]
        if op == "exsum":
            ms = timeit(lambda: h.exsum_async(n, a, 1, 0, fpe, bool(ee)))
            gbs = n * 8 / ms / 1e6
        else:
            ms = timeit(lambda: h.exdot_async(n, a, 1, 0, b, 1, 0, fpe, bool(ee)))
            gbs = n * 16 / ms / 1e6
        row.append(f"{op[2:]}:{kind}:{fpe}{'e' if ee else ''}={gbs:.0f}")
    print(f"adaptive={pf:2d} " + " ".join(row), flush=True)
