"""Run ONE bench configuration a few times (for ncu captures and quick timing), on bench.py's own data.

    python scripts/prof_case.py op dist fpe ee log2n [reps]
      op:   exsum | exdot | gemvN | gemvT
      dist: loguniform | loguniform_signed | naive | illcond | cancel (exdot: the cancelling ill-conditioned pair) | narrow (gemv)
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import exblas_b200 as xb
from exblas_b200 import common as cm
import bench

op, dist, fpe, ee, lg = sys.argv[1], sys.argv[2], int(sys.argv[3]), bool(int(sys.argv[4])), int(sys.argv[5])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
dev = torch.device("cuda:0")
h = xb.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream)
n = 1 << lg
if op.startswith("gemv"):
    m = 1 << (lg // 2)
    A = torch.empty(m * m, dtype=torch.float64, device=dev)
    for lo in range(0, m * m, 1 << 27):
        hi = min(m * m, lo + (1 << 27))
        A[lo:hi] = cm.init_fpuniform(m * m, 10, 5, seed=1, neg_ratio=2, lo=lo, hi=hi, device=dev)
    x = cm.init_fpuniform(m, 10, 5, seed=2, neg_ratio=2, device=dev)
    y = torch.zeros(m, dtype=torch.float64, device=dev)
    run = lambda: xb.exgemv(op[-1], m, m, 1.0, A, m, 0, x, 1, 0, 0.0, y, 1, 0, fpe, ee, handle=h, sync=False)
    nbytes = (m * m + 2 * m) * 8
else:
    if dist == "cancel":
        a, b = cm.cancelling_pair(n, "dot", seed=7, device=dev)
    else:
        a = bench.gen_sliced(dist, n, 0, n, 1, dev)
        b = bench.gen_sliced(dist, n, 0, n, 2, dev) if op == "exdot" else None
    run = (lambda: h.exsum_async(n, a, 1, 0, fpe, ee)) if op == "exsum" else (lambda: h.exdot_async(n, a, 1, 0, b, 1, 0, fpe, ee))
    nbytes = n * (16 if op == "exdot" else 8)
run(); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(reps): run()
e1.record(s); e1.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{op} {dist} fpe={fpe} ee={int(ee)} n=2^{lg}: {ms:.3f} ms {nbytes / ms / 1e6:.1f} GB/s kernel={h.last_kernel()}")
