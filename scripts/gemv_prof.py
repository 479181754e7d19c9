"""Development helper: one ExGEMV configuration a few times (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exblas_b200 as xb
m = n = int(sys.argv[1]); fpe = int(sys.argv[2]); ee = bool(int(sys.argv[3]))
h = xb.Handle(0)
A = torch.full((m * n,), 1.1, dtype=torch.float64, device="cuda")
x = torch.full((n,), 1.1, dtype=torch.float64, device="cuda")
y = torch.zeros(m, dtype=torch.float64, device="cuda")
for _ in range(3):
    xb.exgemv("N", m, n, 1.0, A, m, 0, x, 1, 0, 0.0, y, 1, 0, fpe, ee, handle=h)
print(float(y[0]))
