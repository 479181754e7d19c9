"""BASELINE config 4: ExSUM size sweep 2^10 .. 2^32 doubles on one GPU (latency-bound to HBM-bound
crossover).  Prints one JSON line per size: device time per call (CUDA events, back-to-back async
calls) and wall time of one synchronous call with a device pointer (launch + fetch)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exblas_b200 as xb
from exblas_b200 import common as cm

dev = torch.device("cuda:0")
h = xb.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream)
max_lg = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nmax = 1 << max_lg
a = torch.empty(nmax, dtype=torch.float64, device=dev)
CH = 1 << 27
for lo in range(0, nmax, CH):          # generate in slices: the generator's temporaries are several x the slice
    hi = min(nmax, lo + CH)
    a[lo:hi] = cm.init_fpuniform(nmax, 664, 332, seed=1, neg_ratio=2, lo=lo, hi=hi, device=dev)
torch.cuda.synchronize()
for lg in range(10, max_lg + 1, 2):
    n = 1 << lg
    row = {"log2n": lg}
    for fpe, ee, tag in [(0, False, "fpe0"), (3, False, "fpe3"), (8, True, "fpe8ee")]:
        reps = 200 if lg <= 20 else (20 if lg <= 26 else 5)
        for _ in range(3): h.exsum_async(n, a, 1, 0, fpe, ee)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps): h.exsum_async(n, a, 1, 0, fpe, ee)
        e1.record(s); e1.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        t0 = time.perf_counter()
        for _ in range(max(1, reps // 4)): v = h.exsum(n, a, 1, 0, fpe, ee)
        sync_us = (time.perf_counter() - t0) * 1e6 / max(1, reps // 4)
        row[tag] = {"device_us": round(us, 2), "GBs": round(n * 8 / us / 1e3, 1), "sync_call_us": round(sync_us, 1)}
    print(json.dumps(row), flush=True)
