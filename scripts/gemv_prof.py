"""Development helper: one ExGEMV configuration a few times (for ncu).
    python scripts/gemv_prof.py m fpe ee [trans [kind]]      kind: naive | narrow | loguniform"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exblas_b200 as xb
from exblas_b200 import common as cm
m = n = int(sys.argv[1]); fpe = int(sys.argv[2]); ee = bool(int(sys.argv[3]))
trans = sys.argv[4] if len(sys.argv) > 4 else "N"
kind = sys.argv[5] if len(sys.argv) > 5 else "naive"
h = xb.Handle(0)
if kind == "naive":
    A = torch.full((m * n,), 1.1, dtype=torch.float64, device="cuda")
    x = torch.full((n,), 1.1, dtype=torch.float64, device="cuda")
else:
    rng, emax = (10, 5) if kind == "narrow" else (664, 332)
    A = torch.empty(m * n, dtype=torch.float64, device="cuda")
    for lo in range(0, m * n, 1 << 27):
        hi = min(m * n, lo + (1 << 27))
        A[lo:hi] = cm.init_fpuniform(m * n, rng, emax, seed=1, neg_ratio=2, lo=lo, hi=hi, device="cuda")
    x = cm.init_fpuniform(n, 10, 5, seed=2, neg_ratio=2, device="cuda")
y = torch.zeros(m, dtype=torch.float64, device="cuda")
for _ in range(3):
    xb.exgemv(trans, m, n, 1.0, A, m, 0, x, 1, 0, 0.0, y, 1, 0, fpe, ee, handle=h)
print(float(y[0]))
