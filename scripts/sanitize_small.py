"""Development helper: a small run that touches every kernel family (for compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import exblas_b200 as xb
from exblas_b200 import common as cm
h = xb.Handle(0)
n = 150_001
a = cm.init_fpuniform(n, 664, 332, seed=1, neg_ratio=2); b = cm.init_ill_cond(n, 1e32, seed=2)
da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
vals = []
for fpe, ee in [(0, False), (3, False), (8, False), (8, True)]:
    vals.append(h.exsum(n, da, 1, 0, fpe, ee))
    vals.append(h.exsum(n - 7, da, 1, 3, fpe, ee))
    vals.append(h.exsum(n // 3, da, 3, 1, fpe, ee))
    vals.append(h.exdot(n, da, 1, 0, db, 1, 0, fpe, ee))
    vals.append(h.exsum(5000, a, 1, 0, fpe, ee))
m, k = 300, 257
A = torch.from_numpy(cm.init_fpuniform(m * k, 100, 50, seed=3, neg_ratio=2)).cuda(); x = torch.from_numpy(cm.init_naive(k)).cuda()
y = torch.zeros(m, dtype=torch.float64, device="cuda")
for fpe, ee in [(0, False), (4, False), (8, True)]:
    xb.exgemv("N", m, k, 1.0, A, m, 0, x, 1, 0, 0.0, y, 1, 0, fpe, ee, handle=h)
    xb.exgemv("N", m, k, 0.5, A, m, 0, x, 1, 0, 2.0, y, 1, 0, fpe, ee, handle=h)
torch.cuda.synchronize()
print("ok", len(set(vals)), float(y[0]))
