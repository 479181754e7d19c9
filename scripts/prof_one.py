"""Development helper: run ONE configuration a few times (for ncu / quick timing).
usage: prof_one.py op kind fpe ee log2n [T] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import exblas_b200 as xb

op, kind, fpe, ee, lg = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
T = int(sys.argv[6]) if len(sys.argv) > 6 else 512
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
n = 1 << lg
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)

def gen(kind):
    if kind == "naive":
        return torch.full((n,), 1.1, dtype=torch.float64, device=dev)
    m = torch.rand(n, dtype=torch.float64, device=dev, generator=g) + 1.0
    if kind == "logu":
        e = torch.randint(-332, 332, (n,), device=dev, generator=g)
        s = torch.randint(0, 2, (n,), device=dev, generator=g).double() * 2 - 1
        return torch.ldexp(m, e) * s
    if kind == "ill":
        e = torch.randint(0, 54, (n,), device=dev, generator=g)
        return torch.ldexp(2 * (m - 1.5), e)
    raise ValueError(kind)

a = gen(kind)
b = gen(kind) if op == "exdot" else None
h = xb.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream)
h.set_option("block_threads", T)
def run():
    if op == "exsum": h.exsum_async(n, a, 1, 0, fpe, bool(ee))
    else: h.exdot_async(n, a, 1, 0, b, 1, 0, fpe, bool(ee))
run(); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(reps): run()
e1.record(s); e1.synchronize()
ms = e0.elapsed_time(e1) / reps
v, l, st = h.fetch()
bytes_ = n * 8 * (2 if op == "exdot" else 1)
print(f"{op} {kind} fpe={fpe} ee={ee} n=2^{lg} T={T}: {ms:.3f} ms {bytes_/ms/1e6:.1f} GB/s value={v!r} status={st}")
