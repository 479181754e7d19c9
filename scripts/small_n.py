"""Development helper: a few small-n launches (for an ncu launch list)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exblas_b200 as xb
from exblas_b200 import common as cm
h = xb.Handle(0)
a = cm.init_fpuniform(1 << 20, 664, 332, seed=1, neg_ratio=2, device="cuda")
torch.cuda.synchronize()
for lg in (10, 13, 15, 16, 18, 20):
    for fpe, ee in ((0, False), (3, False), (8, True)):
        for _ in range(3):
            h.exsum_async(1 << lg, a, 1, 0, fpe, ee)
print(h.fetch()[0])
