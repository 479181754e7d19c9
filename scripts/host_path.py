"""Host-pointer path (what blas1.hpp callers use): GB/s of exsum() on PAGEABLE and PINNED host vectors, for several
copy-thread counts and chunk sizes.   python scripts/host_path.py [log2n]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import exblas_b200 as xb
from exblas_b200 import common as cm
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 28
n = 1 << lg
h = xb.Handle(0)
a = np.empty(n, dtype=np.float64)
for lo in range(0, n, 1 << 24):
    a[lo:lo + (1 << 24)] = cm.init_fpuniform(n, 664, 332, seed=1, neg_ratio=1, lo=lo, hi=min(n, lo + (1 << 24)))
pin = torch.from_numpy(a).pin_memory()

def rate(src, reps=3):
    h.exsum(n, src, 1, 0, 3, False)
    t0 = time.perf_counter()
    for _ in range(reps): v = h.exsum(n, src, 1, 0, 3, False)
    return round(n * 8 * reps / (time.perf_counter() - t0) / 1e9, 2), v

print(json.dumps({"cores": os.cpu_count(), "log2n": lg}))
r, v0 = rate(pin)
print(json.dumps({"pinned_GBs": r}), flush=True)
for threads in (1, 2, 4, 6, 8, 12, 16):
    for chunk in (1 << 20, 1 << 21, 1 << 22, 1 << 23):
        if threads == 1 and chunk != 1 << 21: continue
        h.set_option("host_threads", threads); h.set_option("pageable_chunk_elems", chunk)
        r, v = rate(a)
        print(json.dumps({"threads": threads, "chunk_log2": chunk.bit_length() - 1, "pageable_GBs": r, "same": v == v0}), flush=True)
