"""Where does a reduction kernel spend its time?  Option "phase_timing" makes thread 0 of every CTA stamp
%globaltimer at the phase boundaries of exblas_reduce_kernel / reduce_finish; this prints, per size, the
median / max over CTAs of each phase (us) and the span from the first CTA's start to the last stamp.

    python scripts/phase_times.py [log2n ...]
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import exblas_b200 as xb
from exblas_b200 import common as cm

NAMES = ["start", "body", "fin_entry", "remainder", "scalar+flush", "rowsum", "global", "normalised", "peers", "published"]
dev = torch.device("cuda:0")
h = xb.Handle(0)
sizes = [int(x) for x in sys.argv[1:]] or [10, 13, 16, 20, 24, 28]
a = cm.init_fpuniform(1 << max(sizes), 664, 332, seed=1, neg_ratio=1, device=dev) if max(sizes) <= 27 else None
if a is None:
    a = torch.empty(1 << max(sizes), dtype=torch.float64, device=dev)
    for lo in range(0, a.numel(), 1 << 27):
        a[lo:lo + (1 << 27)] = cm.init_fpuniform(a.numel(), 664, 332, seed=1, neg_ratio=1, lo=lo, hi=lo + (1 << 27), device=dev)
h.set_option("phase_timing", 1)
for lg in sizes:
    n = 1 << lg
    for fpe, ee in ((0, False), (3, False), (8, True)):
        for _ in range(3):
            h.exsum_async(n, a, 1, 0, fpe, ee)
        h.fetch()
        h.phase_times()
        h.exsum_async(n, a, 1, 0, fpe, ee)
        h.fetch()
        t = h.phase_times().astype(np.int64)
        t0 = t[:, 0].min()
        row = {"log2n": lg, "fpe": fpe, "ee": ee, "ctas": int(t.shape[0])}
        ph = {}
        for k in range(1, 10):
            col = t[:, k]
            ok = col > 0
            if not ok.any():
                continue
            prev = np.maximum.reduce([t[:, j] for j in range(k)])        # last stamp before k in the same CTA
            d = (col - prev)[ok] / 1e3
            ph[NAMES[k]] = {"med": round(float(np.median(d)), 2), "max": round(float(d.max()), 2), "n": int(ok.sum())}
        row["phases_us"] = ph
        row["start_skew_us"] = round(float(t[:, 0].max() - t0) / 1e3, 2)
        row["kernel_span_us"] = round(float(t.max() - t0) / 1e3, 2)
        print(json.dumps(row), flush=True)
