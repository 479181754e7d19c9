"""Development helper: A/B library builds under SUSTAINED load (power-cap clocks), alternating rounds.

    python scripts/ab_sustained.py [--rounds 2] [--seconds 1.0] [--cfg sum:logupos:3,...] lib_a.so lib_b.so ...

"default" = the in-tree library.  Every (library, round) runs in its own process (EXBLAS_B200_LIB), generates the same
2^30-element vectors and times each configuration for `seconds` of back-to-back launches (CUDA events).  Prints one
JSON line per (library, round) with GB/s per configuration; the parity of every configuration's limbs against the
default library's first round is checked too (`same_bits`)."""
import argparse, hashlib, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEFAULT_CFG = "sum:logupos:3,sum:logupos:8,sum:logu:3,sum:naive:3,sum:naive:4,sum:naive:8,sum:naive:8e,dot:ill:0,dot:ill:3,dot:ill:8e"


def child(cfgs, seconds, log2n):
    import torch
    import exblas_b200 as xb
    from exblas_b200 import common as cm
    dev = torch.device("cuda:0")
    h = xb.Handle(0)
    s = torch.cuda.Stream(); torch.cuda.set_stream(s); h.set_stream(s.cuda_stream)
    n = 1 << log2n
    cache = {}

    def sliced(f):
        out = torch.empty(n, dtype=torch.float64, device=dev)
        for lo in range(0, n, 1 << 26):
            out[lo:lo + (1 << 26)] = f(lo, min(n, lo + (1 << 26)))
        return out

    def data(kind):
        if kind not in cache:
            cache.clear(); torch.cuda.empty_cache()
            if kind == "logupos": cache[kind] = (sliced(lambda lo, hi: cm.init_fpuniform(n, 664, 332, seed=1, neg_ratio=1, lo=lo, hi=hi, device=dev)), None)
            elif kind == "logu": cache[kind] = (sliced(lambda lo, hi: cm.init_fpuniform(n, 664, 332, seed=1, neg_ratio=2, lo=lo, hi=hi, device=dev)), None)
            elif kind == "naive": cache[kind] = (cm.init_naive(n, device=dev), None)
            elif kind == "ill": cache[kind] = cm.cancelling_pair(n, "dot", seed=7, device=dev)
            elif kind == "narrow": cache[kind] = (sliced(lambda lo, hi: cm.init_fpuniform(n, 10, 5, seed=1, neg_ratio=2, lo=lo, hi=hi, device=dev)),
                                                  sliced(lambda lo, hi: cm.init_fpuniform(n, 10, 5, seed=2, neg_ratio=2, lo=lo, hi=hi, device=dev)))
            elif kind == "illsum": cache[kind] = (sliced(lambda lo, hi: cm.init_ill_cond(n, 1e32, seed=1, lo=lo, hi=hi, device=dev)), None)
        return cache[kind]

    res = {}
    for c in cfgs:
        op, kind, f = c.split(":")
        ee = f.endswith("e"); fpe = int(f.rstrip("e"))
        a, b = data(kind)
        if op == "dot" and b is None:
            b = a
        fn = (lambda: h.exsum_async(n, a, 1, 0, fpe, ee)) if op == "sum" else (lambda: h.exdot_async(n, a, 1, 0, b, 1, 0, fpe, ee))
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); fn(); e1.record(s); e1.synchronize()
        reps = max(3, int(seconds * 1e3 / e0.elapsed_time(e1)))
        e0.record(s)
        for _ in range(reps): fn()
        e1.record(s); e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
        v, l, st = h.fetch()
        sha = hashlib.sha256(l.tobytes() + repr(v).encode()).hexdigest()[:12]
        res[c] = [round(n * (16 if op == "dot" else 8) / (ms * 1e-3) / 1e9, 1), sha]
    print(json.dumps(res), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--seconds", type=float, default=1.0)
    ap.add_argument("--log2n", type=int, default=30)
    ap.add_argument("--cfg", default=DEFAULT_CFG)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("libs", nargs="*")
    args = ap.parse_args()
    cfgs = args.cfg.split(",")
    if args.child:
        return child(cfgs, args.seconds, args.log2n)
    libs = [("default", None)] + [(os.path.basename(p), os.path.abspath(p)) for p in args.libs]
    ref = None
    for rnd in range(args.rounds):
        for name, path in libs:
            env = dict(os.environ)
            if path:
                env["EXBLAS_B200_LIB"] = path
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--seconds", str(args.seconds), "--log2n", str(args.log2n),
                                "--cfg", args.cfg], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else None
            if not line:
                print(json.dumps({"lib": name, "round": rnd, "error": p.stderr[-400:]}), flush=True)
                continue
            res = json.loads(line)
            if ref is None:
                ref = {c: v[1] for c, v in res.items()}
            print(json.dumps({"lib": name, "round": rnd, "GBs": {c: v[0] for c, v in res.items()},
                              "same_bits": all(ref.get(c) == v[1] for c, v in res.items())}), flush=True)


if __name__ == "__main__":
    main()
