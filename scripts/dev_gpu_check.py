"""Development helper: quick parity + timing sweep on a real GPU (not part of the test-suite)."""
import math, sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import exblas_b200 as xb
from oracle.oracle import Oracle

O = Oracle()
dev = torch.device("cuda:0")
h = xb.Handle(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
assert stream.cuda_stream != 0
h.set_stream(stream.cuda_stream)
QUICK = len(sys.argv) > 1 and sys.argv[1] == "quick"
rng = np.random.default_rng(7)

def gen(kind, n):
    if kind == "naive": return np.full(n, 1.1)
    if kind == "logu": return np.ldexp(rng.uniform(1, 2, n), rng.integers(-332, 332, n)) * rng.choice([-1.0, 1.0], n)
    if kind == "logupos": return np.ldexp(rng.uniform(1, 2, n), rng.integers(-332, 332, n))
    if kind == "ill": return (2 * rng.random(n) - 1) * np.ldexp(1.0, rng.integers(0, 54, n))
    raise ValueError(kind)

fails = 0
for kind in ["naive", "logu", "logupos", "ill"]:
    for n in ([1000, 100003] if QUICK else [0, 1, 3, 5, 31, 1000, 8191, 8192, 8193, 100003, 1 << 20, (1 << 21) + 5]):
        a = gen(kind, n)
        ro, lo = O.exsum(a, fpe=0)
        rx, _ = O.exsum(a, fpe=0, round_mode=1)
        d = torch.from_numpy(a).to(dev)
        for fpe, ee in [(0, 0), (2, 0), (3, 0), (4, 0), (8, 0), (4, 1), (6, 1), (8, 1)]:
            for src in (d, a):
                v, l = h.exsum(n, src, 1, 0, fpe, bool(ee), want_limbs=True)
                if v != ro or not (l == lo).all():
                    fails += 1
                    print("MISMATCH exsum", kind, n, fpe, ee, type(src).__name__, v, ro)
            v1 = h.exsum(n, d, 1, 0, fpe, bool(ee), round_mode=1)
            if v1 != rx:
                fails += 1; print("MISMATCH exact", kind, n, fpe, ee, v1, rx)
        # misaligned / strided
        if n > 10:
            for off, inc in [(1, 1), (3, 1), (0, 2), (5, 3)]:
                m = (n - off + inc - 1) // inc
                ro2, lo2 = O.exsum(a, inca=inc, offset=off, n=m)
                v, l = h.exsum(m, d, inc, off, 4, False, want_limbs=True)
                if v != ro2 or not (l == lo2).all():
                    fails += 1; print("MISMATCH strided", kind, n, off, inc)
        # exdot
        b = gen(kind, n)
        rd, ld = O.exdot(a, b, fpe=0)
        db = torch.from_numpy(b).to(dev)
        for fpe, ee in [(0, 0), (3, 0), (4, 0), (8, 0), (4, 1), (8, 1)]:
            v, l = h.exdot(n, d, 1, 0, db, 1, 0, fpe, bool(ee), want_limbs=True)
            if n == 0: continue
            if v != rd or not (l == ld).all():
                fails += 1; print("MISMATCH exdot", kind, n, fpe, ee, v, rd)
print("parity fails:", fails, "status", h.last_status())

# timing
n = 1 << 28
res = []
for kind in ["naive", "logu", "ill"]:
    a = gen(kind, n)
    d = torch.from_numpy(a).to(dev)
    del a
    for T, AD in [(512, 1), (512, 0)]:
        h.set_option("block_threads", T); h.set_option("adaptive", AD)
        for fpe, ee in [(0, 0), (2, 0), (3, 0), (4, 0), (8, 0), (4, 1), (8, 1)]:
            for _ in range(2): h.exsum_async(n, d, 1, 0, fpe, bool(ee))
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            K = 5
            for _ in range(K): h.exsum_async(n, d, 1, 0, fpe, bool(ee))
            e1.record(stream); e1.synchronize()
            ms = e0.elapsed_time(e1) / K
            v, _, st = h.fetch()
            gbs = n * 8 / ms / 1e6
            res.append((kind, T, fpe, ee, ms, gbs))
            print(f"exsum {kind:6s} T={T} ad={AD} fpe={fpe} ee={ee}: {ms:8.3f} ms  {gbs:8.1f} GB/s  v={v!r} st={st}", flush=True)
    if kind != "naive":
        b = torch.from_numpy(gen(kind, n)).to(dev)
        h.set_option("block_threads", 512); h.set_option("adaptive", 1)
        for fpe, ee in [(0, 0), (3, 0), (4, 0), (8, 0), (4, 1), (8, 1)]:
            for _ in range(2): h.exdot_async(n, d, 1, 0, b, 1, 0, fpe, bool(ee))
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            K = 5
            for _ in range(K): h.exdot_async(n, d, 1, 0, b, 1, 0, fpe, bool(ee))
            e1.record(stream); e1.synchronize()
            ms = e0.elapsed_time(e1) / K
            v, _, st = h.fetch()
            print(f"exdot {kind:6s} fpe={fpe} ee={ee}: {ms:8.3f} ms  {2*n*8/ms/1e6:8.1f} GB/s v={v!r} st={st}", flush=True)
        del b
    del d
