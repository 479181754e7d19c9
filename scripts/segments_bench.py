"""Batched ExSUM / ExDOT / CSR SpMV throughput on one B200 (CUDA events): total 2^27 elements cut into
equal segments of length L.  Reports segments/s and GB/s (8 B per element for sums, 16 for dots)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exblas_b200 as xb
from exblas_b200 import common as cm
total = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 27)
dev = torch.device("cuda:0")
h = xb.Handle(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream)
lib = h.lib
for kind, rng in (("narrow", (10, 5)), ("loguniform", (664, 332))):
    a = cm.init_fpuniform(total, rng[0], rng[1], seed=1, neg_ratio=2, device=dev)
    b = cm.init_fpuniform(total, 10, 5, seed=2, neg_ratio=2, device=dev)
    for L in (16, 64, 256, 4096, 1 << 20):
        nseg = total // L
        seg = torch.arange(0, nseg + 1, dtype=torch.int64, device=dev) * L
        out = torch.empty(nseg, dtype=torch.float64, device=dev)
        for op in ("sum", "dot"):
            def run():
                if op == "sum":
                    rc = lib.exblas_b200_exsum_segments(h._h, a.data_ptr(), seg.data_ptr(), nseg, 0, 0, 0, out.data_ptr(), None)
                else:
                    rc = lib.exblas_b200_exdot_segments(h._h, a.data_ptr(), b.data_ptr(), None, 0, seg.data_ptr(), nseg, 0, 0, 0, out.data_ptr(), None)
                assert rc == 0
            for _ in range(2): run()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(s)
            K = 5
            for _ in range(K): run()
            e1.record(s); e1.synchronize()
            ms = e0.elapsed_time(e1) / K
            bytes_per = 8 if op == "sum" else 16
            print(json.dumps({"op": "ex" + op + "_segments", "data": kind, "segment_len": L, "nseg": nseg, "ms": round(ms, 3),
                              "Msegments_per_s": round(nseg / ms / 1e3, 2), "GBs": round(total * bytes_per / ms / 1e6, 1)}), flush=True)
# CSR SpMV: 2^20 rows x 32 nnz
nrows, nnz_row, ncols = 1 << 20, 32, 1 << 20
vals = cm.init_fpuniform(nrows * nnz_row, 10, 5, seed=3, neg_ratio=2, device=dev)
x = cm.init_fpuniform(ncols, 10, 5, seed=4, neg_ratio=2, device=dev)
colidx = torch.randint(0, ncols, (nrows * nnz_row,), dtype=torch.int32, device=dev)
rowptr = torch.arange(0, nrows + 1, dtype=torch.int64, device=dev) * nnz_row
y = torch.empty(nrows, dtype=torch.float64, device=dev)
def spmv():
    assert lib.exblas_b200_exdot_segments(h._h, vals.data_ptr(), x.data_ptr(), colidx.data_ptr(), ncols, rowptr.data_ptr(), nrows, 0, 0, 0, y.data_ptr(), None) == 0
for _ in range(2): spmv()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(5): spmv()
e1.record(s); e1.synchronize()
ms = e0.elapsed_time(e1) / 5
print(json.dumps({"op": "exact CSR SpMV", "rows": nrows, "nnz_per_row": nnz_row, "ms": round(ms, 3), "Gnnz_per_s": round(nrows * nnz_row / ms / 1e6, 2)}))
