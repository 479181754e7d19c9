import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import exblas_b200 as xb
from exblas_b200 import common as cm
from oracle.oracle import Oracle
O = Oracle(); h = xb.Handle(0)
def run(lengths, kind="wide"):
    seg = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    total = int(seg[-1])
    a = cm.init_fpuniform(total, 664, 332, seed=21, neg_ratio=2) if kind == "wide" else cm.init_fpuniform(total, 664, 332, seed=21, neg_ratio=1)
    got = h.exsum_segments(torch.from_numpy(a).cuda(), torch.from_numpy(seg).cuda()).cpu().numpy()
    bad = []
    for i in list(range(min(3, len(lengths)))) + [len(lengths) - 1]:
        w = O.exsum(a[seg[i]:seg[i+1]], fpe=0)[0] if seg[i+1] > seg[i] else 0.0
        if np.float64(got[i]).view(np.uint64) != np.float64(w).view(np.uint64): bad.append((i, got[i], w))
    print(len(lengths), lengths[0], kind, "R~", (total + 2367) // 2368, "bad:", bad, flush=True)
for L in (100_000, 1_000_000, 2_000_000, 2_400_000, 2_600_000, 2_800_000, 3_000_000, 4_000_000):
    run([L] + [7] * 500)
run([3_000_000]); run([3_000_000, 7]); run([3_000_000] + [7] * 500, "pos"); run([6_000_000]); run([6_000_000], "pos")
