"""Development helper: ExGEMV 32768 x 32768 A/B of library builds (EXBLAS_B200_LIB), alternating rounds, one process per
(library, round); prints GB/s per (trans, data, shape) and a hash of y (must agree between libraries).

    python scripts/gemv_ab.py [--rounds 2] [--cfg T:narrow:2,N:narrow:1,...] [--opt name=value ...] lib_a.so ...
cfg entries: trans:data:shape (shape = option gemv_t_shape / gemv_n_shape; data = narrow | naive | loguniform)"""
import argparse, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(cfgs, m, opts):
    import torch
    import exblas_b200 as xb
    from exblas_b200 import common as cm
    n = m
    dev = torch.device("cuda:0")
    h = xb.Handle(0)
    s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream)
    for o in opts:
        k, v = o.split("="); h.set_option(k, int(v))
    A = torch.empty(m * n, dtype=torch.float64, device=dev)
    cur = None
    res = {}
    for c in cfgs:
        trans, kind, shape = c.split(":")
        if kind != cur:
            CH = 1 << 27
            for lo in range(0, m * n, CH):
                hi = min(m * n, lo + CH)
                if kind == "naive": A[lo:hi] = 1.1
                elif kind == "loguniform": A[lo:hi] = cm.init_fpuniform(m * n, 664, 332, seed=1, neg_ratio=2, lo=lo, hi=hi, device=dev)
                else: A[lo:hi] = cm.init_fpuniform(m * n, 10, 5, seed=1, neg_ratio=2, lo=lo, hi=hi, device=dev)
            x = torch.full((n,), 1.1, dtype=torch.float64, device=dev) if kind == "naive" else cm.init_fpuniform(n, 10, 5, seed=2, neg_ratio=2, device=dev)
            cur = kind
        h.set_option("gemv_t_shape" if trans == "T" else "gemv_n_shape", int(shape))
        y = torch.zeros(m, dtype=torch.float64, device=dev)
        run = lambda: xb.exgemv(trans, m, n, 1.0, A, m, 0, x, 1, 0, 0.0, y, 1, 0, 0, False, handle=h, sync=False)
        for _ in range(3): run()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        K = 10
        e0.record(s)
        for _ in range(K): run()
        e1.record(s); e1.synchronize()
        ms = e0.elapsed_time(e1) / K
        res[c] = [round((m * n + m + n) * 8 / ms / 1e6, 1), hashlib.sha256(y.cpu().numpy().tobytes()).hexdigest()[:10], h.last_status()]
    print(json.dumps(res), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--m", type=int, default=32768)
    ap.add_argument("--cfg", default="T:narrow:2,N:narrow:1,T:naive:2,N:naive:1,T:loguniform:2,N:loguniform:1")
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--child", action="store_true")
    ap.add_argument("libs", nargs="*")
    a = ap.parse_args()
    if a.child:
        return child(a.cfg.split(","), a.m, a.opt)
    libs = [("default", None)] + [(os.path.basename(p), os.path.abspath(p)) for p in a.libs]
    ref = None
    for rnd in range(a.rounds):
        for name, path in libs:
            env = dict(os.environ)
            if path: env["EXBLAS_B200_LIB"] = path
            cmd = [sys.executable, os.path.abspath(__file__), "--child", "--m", str(a.m), "--cfg", a.cfg] + [x for o in a.opt for x in ("--opt", o)]
            p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else None
            if not line:
                print(json.dumps({"lib": name, "round": rnd, "error": p.stderr[-600:]}), flush=True); continue
            res = json.loads(line)
            if ref is None: ref = {c: v[1] for c, v in res.items()}
            print(json.dumps({"lib": name, "round": rnd, "GBs": {c: v[0] for c, v in res.items()}, "status": sorted({v[2] for v in res.values()}),
                              "same_bits": all(ref.get(c) == v[1] for c, v in res.items())}), flush=True)


if __name__ == "__main__":
    main()
