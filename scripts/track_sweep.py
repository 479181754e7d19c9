"""Development helper: device time per ExSUM reduction (CUDA-graph replay) with option track_max_elems = 0 (the kernel that
fpe selects) against track_max_elems = 2^24 (every size of the sweep takes the superaccumulator-only kernel without its
unrolled body: all summands pass through the row-tracking loops and the merge sums only the touched limb rows).
    python scripts/track_sweep.py [loguniform | narrow]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exblas_b200 as xb
from exblas_b200 import common as cm
dev = torch.device("cuda:0")
h = xb.Handle(0)
kind = sys.argv[1] if len(sys.argv) > 1 else "loguniform"
a = cm.init_fpuniform(1 << 24, 664, 332, seed=1, neg_ratio=1, device=dev) if kind == "loguniform" else cm.init_fpuniform(1 << 24, 10, 5, seed=1, neg_ratio=2, device=dev)
s = torch.cuda.Stream()
h.set_stream(s.cuda_stream)
K = 20
for lg in (14, 16, 18, 19, 20, 21, 22):
    n = 1 << lg
    row = {"log2n": lg, "data": kind}
    ref = None
    for track in (0, 1 << 24):
        h.set_option("track_max_elems", track)
        for fpe in (0, 3, 8):
            with torch.cuda.stream(s):
                for _ in range(3): h.exsum_async(n, a, 1, 0, fpe, False)
                s.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    for _ in range(K): h.exsum_async(n, a, 1, 0, fpe, False)
                g.replay(); s.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                for _ in range(10): g.replay()
                e1.record(s); e1.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (10 * K)
            v, limbs, st = h.fetch()
            if ref is None: ref = (v, limbs.tobytes())
            row[f"track{int(track > 0)}_fpe{fpe}"] = [round(us, 2), (v, limbs.tobytes()) == ref]
    print(json.dumps(row), flush=True)
