#!/usr/bin/env python
"""bench.py -- ExSUM / ExDOT throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--op exsum|exdot] [--dist loguniform|loguniform_signed|naive|illcond] [--log2n 30] [--fpe 3,4,8]
    torchrun ... bench.py --gpus N ...          (N > 1: one rank per GPU)

Workload (default = BASELINE.json configs[1]): ExSUM of ONE vector of 2^30 doubles (8 GiB), log-uniform
1e-100..1e100 (the reference's init_fpuniform(n, 664, 332), positive values), FPE sizes 3, 4 and 8.  One STEP = one
reduction per FPE size over the same resident vector, i.e. 3 kernel launches and 3 x 8 GiB of algorithmic traffic.
`value` = algorithmic bytes / device time (CUDA events on the launching stream, inputs resident in HBM; 8 GiB >> 126 MB
L2, so every pass streams from DRAM).

N > 1 is STRONG scaling (BASELINE configs[2]/[3], the reference's MPI path cpu ExSUM.cpp:33-65,266-273): the SAME
2^30-element vector (same seed) is sharded into N contiguous ranges, every rank reduces its range on its GPU and the
ranks combine their limbs exactly -- by default inside the closing kernel over NVLink peer memory (--collective
fused), or with ncclAllReduce (--collective nccl).  `result.limbs_sha` is the SHA-256 of the 39 normalised limbs +
the rounded value: it is the same string for N = 1, 2, 4, 8, which is the product.  The weak-scaling figure (2^30
elements PER GPU) is reported in `extras.weak_scaling`.

`e2e` = the same step through the synchronous C-ABI entry point with HOST buffers (pinned, and pageable as
`e2e.pageable`): H2D copies and the D2H read of the result are inside the timed region.

`roofline` is the slower of two ceilings, both measured in this run: HBM (MEASURED_PEAKS.json copy bandwidth; the
read-only stream ceiling is reported beside it) and the FP64 pipe (DADD lane-instructions/s x the variant's FP64
instructions per element).

--impl reference times the reference's own CPU ExSUM (oracle/_ref, unmodified sources, OpenMP over all host cores;
the oracle port if that prebuilt library is absent) on a bounded sample of the same workload per step.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) off it
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

METRIC = "ExSUM/ExDOT GB/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--op", default="exsum", choices=["exsum", "exdot"])
    ap.add_argument("--dist", default="loguniform", choices=["loguniform", "loguniform_signed", "naive", "illcond"])
    ap.add_argument("--log2n", type=int, default=30, help="log2 of the TOTAL vector length (sharded over the GPUs)")
    ap.add_argument("--fpe", default="3,4,8")
    ap.add_argument("--early-exit", type=int, default=0)
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl"],
                    help="N > 1: limb exchange inside the closing kernel over peer memory (fused) or ncclAllReduce")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the side measurements (ExDOT, variants, latency, ExGEMV ...)")
    ap.add_argument("--cpu-log2n", type=int, default=0, help="sample size of the CPU legs (0 = sized to the time budget)")
    return ap.parse_args()


def gen(dist, n_total, lo, hi, seed, device):
    from exblas_b200 import common as cm
    if dist == "naive":
        return cm.init_naive(n_total, lo=lo, hi=hi, device=device)
    if dist == "loguniform":            # the reference's generator: positive values (common.cpp:18-33, neg_ratio = 1)
        return cm.init_fpuniform(n_total, 664, 332, seed=seed, neg_ratio=1, lo=lo, hi=hi, device=device)
    if dist == "loguniform_signed":     # same magnitudes, random sign
        return cm.init_fpuniform(n_total, 664, 332, seed=seed, neg_ratio=2, lo=lo, hi=hi, device=device)
    return cm.init_ill_cond(n_total, 1e32, seed=seed, lo=lo, hi=hi, device=device)


def gen_sliced(dist, n_total, lo, hi, seed, device, out=None):
    """the same vector, generated in slices (the generators' temporaries are several times the slice)"""
    import numpy as np
    step = 1 << 26
    if device is None:
        out = np.empty(hi - lo, dtype=np.float64) if out is None else out
    else:
        import torch
        out = torch.empty(hi - lo, dtype=torch.float64, device=device) if out is None else out
    for s in range(lo, hi, step):
        e = min(hi, s + step)
        out[s - lo:e - lo] = gen(dist, n_total, s, e, seed, device)
    return out


DIST_TEXT = {"loguniform": "log-uniform 1e-100..1e100 (the reference's init_fpuniform(n,664,332): positive values)",
             "loguniform_signed": "log-uniform 1e-100..1e100 (init_fpuniform(n,664,332) magnitudes, random sign)",
             "naive": "all 1.1 (init_naive)", "illcond": "init_ill_cond(n, 1e32)"}


def config_of(args, world):
    """identical in both arms (the driver compares them)"""
    return {"workload": f"{args.op.upper()} of ONE vector of n=2^{args.log2n} doubles, {DIST_TEXT[args.dist]}, "
                        f"FPE sizes {args.fpe}{' early-exit' if args.early_exit else ''}; one step = one reduction per FPE size",
            "n_total": 1 << args.log2n, "fpe": args.fpe, "early_exit": int(bool(args.early_exit)),
            "l2": "inputs (8 GiB per pass at 2^30) are far larger than the 126 MB L2; no flush needed",
            "parallelism": f"shard{world}" if world > 1 else "single"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write, burst)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(op, dist, fpe, ee):
    """dram bytes per launch from the committed `ncu --set full` capture of exactly this variant, else None
    (profiles/traffic.json, keys "op|dist|fpe|ee|log2n")."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))
        return t.get(f"{op}|{dist}|{fpe}|{int(bool(ee))}|30")
    except Exception:
        return None


def fp64_per_elem(op, dist, fpe, ee):
    """FP64-pipe instructions per element of the path this variant takes on this data."""
    dot = op == "exdot"
    f = fpe if fpe >= (3 if dot else 2) else 0
    if ee and f:
        f = 4 if f <= 4 else (6 if f <= 6 else 8)
    wide = dist.startswith("loguniform")
    if f == 0 or wide:                  # direct deposits (superaccumulator-only, or thrash bypass on wide-range data)
        return (2 + 8) if dot else 4, "direct deposits: 4 DADD per summand (magic-constant digit split)" + (" + DMUL/DFMA" if dot else "")
    if ee:                              # narrow data: the walk stops after ~2 levels
        return (2 + 12 + 6) if dot else 12, "early exit, ~2 levels visited on narrow-range data (6 DADD per level)"
    return (2 + 6 * f + 18) if dot else 6 * f, f"{f} levels x 6 DADD (Knuth TwoSum)" + (" + TwoProd + error term through 3 levels" if dot else "")


def roofline_of(op, dist, fpe, ee, n, ms, peak, peak_src, r_fp64, read_gbs, kernel=None):
    """slower-of-two roofline for ONE launch over n elements that took `ms`."""
    bpe = 16 if op == "exdot" else 8
    achieved = n * bpe / (ms * 1e-3) / 1e9
    instr, why = fp64_per_elem(op, dist, fpe, ee)
    t_hbm = n * bpe / (peak * 1e9)
    t_fp = n * instr / r_fp64 if r_fp64 else 0.0
    bound = "fp64" if t_fp > t_hbm else "hbm"
    t_roof = max(t_hbm, t_fp)
    rl = {"bound": bound, "achieved": round(achieved, 1),
          "peak": round(peak if bound == "hbm" else n * bpe / t_fp / 1e9, 1), "unit": "GB/s",
          "frac": round(t_roof / (ms * 1e-3), 4), "traffic": ncu_traffic(op, dist, fpe, ee),
          "fp64_instr_per_elem": instr, "fp64_path": why,
          "hbm_peak_GBs": peak, "fp64_rate_per_s": r_fp64,
          "frac_of_hbm_copy_peak": round(achieved / peak, 4),
          "frac_of_read_stream": round(achieved / read_gbs, 4) if read_gbs else None}
    if bound == "fp64":
        rl["note"] = "FP64-pipe bound: peak = the GB/s at which n x fp64_instr_per_elem saturates the measured DADD rate"
    if kernel:
        rl["kernel"] = kernel
    return rl


def limbs_sha(value, limbs):
    import numpy as np
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(limbs, dtype=np.int64).tobytes())
    h.update(np.float64(value).tobytes())
    return h.hexdigest()[:32]


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own CPU ExSUM on the host cores
# ------------------------------------------------------------------------------------------------
class CpuExsum:
    """the reference CPU library (kind "reference") or, if its prebuilt .so is absent, the oracle port"""

    def __init__(self, op):
        from oracle.oracle import Oracle, Reference
        self.op = op
        if op == "exsum" and Reference.available():
            self.impl = Reference()
            self.kind, self.cores = "reference", self.impl.use_all_cores()
        else:
            self.impl = Oracle()
            self.kind, self.cores = "port", (self.impl.use_all_cores() if op == "exsum" else 1)

    def run(self, a, b, fpe, ee):
        if self.kind == "reference":
            return self.impl.exsum(a, fpe=fpe, early_exit=ee, parallel=True)
        if self.op == "exsum":
            return self.impl.exsum_parallel(a, fpe=fpe, early_exit=ee)
        return self.impl.exdot(a, b, fpe=fpe, early_exit=ee)[0]


def cpu_sample_log2n(args, cpu, fpes, steps_total, budget_s):
    """largest power-of-two sample (<= the workload) whose steps_total steps fit the time budget"""
    import numpy as np
    if args.cpu_log2n:
        return min(args.cpu_log2n, args.log2n)
    probe = 1 << min(22, args.log2n)
    a = np.ascontiguousarray(gen(args.dist, 1 << args.log2n, 0, probe, 1, None))
    b = np.ascontiguousarray(gen(args.dist, 1 << args.log2n, 0, probe, 2, None)) if args.op == "exdot" else None
    cpu.run(a, b, fpes[0], bool(args.early_exit))
    t0 = time.perf_counter()
    for f in fpes:
        cpu.run(a, b, f, bool(args.early_exit))
    per_elem = (time.perf_counter() - t0) / probe            # seconds per element and step
    lg = args.log2n
    while lg > 20 and per_elem * (1 << lg) * steps_total > budget_s:
        lg -= 1
    return min(lg, 29)                                       # (2^29 doubles = 4 GiB of host memory at most)


def cpu_run(args, fpes, log2n, steps, warmup, cpu=None):
    """-> (GB/s, seconds per step, description).  `steps` timed steps after `warmup` untimed ones."""
    import numpy as np
    cpu = cpu or CpuExsum(args.op)
    n = 1 << log2n
    a = gen_sliced(args.dist, 1 << args.log2n, 0, n, 1, None)
    b = gen_sliced(args.dist, 1 << args.log2n, 0, n, 2, None) if args.op == "exdot" else None
    ee = bool(args.early_exit)
    for _ in range(max(warmup, 1)):
        for f in fpes:
            cpu.run(a, b, f, ee)
    best = {}
    t_all = []
    for _ in range(steps):
        t0 = time.perf_counter()
        for f in fpes:
            t1 = time.perf_counter()
            cpu.run(a, b, f, ee)
            best[f] = min(best.get(f, 1e30), time.perf_counter() - t1)
        t_all.append(time.perf_counter() - t0)
    bpe = 16 if args.op == "exdot" else 8
    t_step = sum(t_all) / len(t_all)
    gbs = n * bpe * len(fpes) / t_step / 1e9
    desc = {"value": round(gbs, 3), "unit": "GB/s", "cores": cpu.cores, "kind": cpu.kind,
            "sample": f"first 2^{log2n} elements of the same vector per step, FPE {','.join(map(str, fpes))}"
                      f"{' early-exit' if ee else ''}, mean of {steps} steps; best per-FPE GB/s: " +
                      ", ".join(f"{f}:{n * bpe / best[f] / 1e9:.2f}" for f in fpes)}
    return gbs, t_step, desc


def cpu_config0(cpu):
    """BASELINE configs[0]: the CPU reference, FPE = 8 with early exit, 2^25 doubles, naive / log-uniform / ill-conditioned
    (tests/test.exsum.cpu.cpp:107-112).  GB/s, best of 3."""
    import numpy as np
    out = {}
    n = 1 << 25
    for dist in ("naive", "loguniform", "illcond"):
        a = np.ascontiguousarray(gen(dist, n, 0, n, 1, None))
        cpu.run(a, None, 8, True)
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            cpu.run(a, None, 8, True)
            best = min(best, time.perf_counter() - t0)
        out[dist] = round(n * 8 / best / 1e9, 2)
    return {"workload": "ExSUM 2^25 doubles, FPE 8 early-exit (tests/test.exsum.cpu.cpp:107-112)", "GBs": out,
            "cores": cpu.cores, "kind": cpu.kind}


def run_reference(args, fpes):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cpu = CpuExsum(args.op)
    warm = max(args.warmup, 1)
    lg = cpu_sample_log2n(args, cpu, fpes, args.steps + warm, 75.0)
    gbs, t_step, desc = cpu_run(args, fpes, lg, args.steps, warm, cpu)
    extras = None
    if args.op == "exsum" and not args.no_extras:
        try:
            extras = {"config0_cpu_reference": cpu_config0(cpu)}
        except Exception as ex:
            extras = {"config0_cpu_reference": repr(ex)}
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gbs, 3), "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": warm, "ms_per_step": round(t_step * 1e3, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_of(args, world),
        "cpu_baseline": desc,
        "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "extras": extras,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, fpes):
    import numpy as np
    import torch
    import exblas_b200 as xb
    from exblas_b200 import common as cm
    from exblas_b200 import dist as xd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    ee = bool(args.early_exit)
    n_total = 1 << args.log2n
    lo, hi = xd.shard_bounds(n_total, rank, world)
    n = hi - lo                                        # this rank's shard
    bpe = 16 if args.op == "exdot" else 8

    h = xb.Handle(local_rank)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)
    red = None
    if world > 1:
        red = xd.DistributedReducer(h)
        red.init_nccl()
        if args.collective == "fused" and not red.init_peer():
            args.collective = "nccl (fused exchange unavailable: CUDA IPC mapping failed)"
            if rank == 0:
                print("bench: falling back to the NCCL limb all-reduce", file=sys.stderr)

    a = gen_sliced(args.dist, n_total, lo, hi, 1, dev)
    b = gen_sliced(args.dist, n_total, lo, hi, 2, dev) if args.op == "exdot" else None
    torch.cuda.synchronize()

    def reduce_async(op, nn, xa, xb_, fpe, e_, rm):
        if op == "exsum":
            h.exsum_async(nn, xa, 1, 0, fpe, e_, rm)
        else:
            h.exdot_async(nn, xa, 1, 0, xb_, 1, 0, fpe, e_, rm)
        if world > 1:
            h.allreduce_async(rm)

    def one(fpe):
        reduce_async(args.op, n, a, b, fpe, ee, xb.ROUND_REFERENCE)

    def step():
        for f in fpes:
            one(f)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, reps):
        """mean device ms per call of fn over `reps` back-to-back calls (max over ranks)"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        e1.synchronize()
        return max_over_ranks(e0.elapsed_time(e1) / reps)

    # ---- ceilings measured in this run (rank-local; idle clocks = burst figures) --------------------
    peak, peak_src = measured_peak()
    try:
        r_fp64 = h.microbench(0)
        read_gbs = h.microbench(1, a)
    except Exception as ex:
        print(f"bench: microbench failed: {ex!r}", file=sys.stderr)
        r_fp64, read_gbs = 1.85e13, None

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    results, kernels = {}, {}
    # per-variant timing (outside the headline region): average kernel duration per FPE size, 3 launches at a time
    per_fpe_ms = {}
    for f in fpes:
        per_fpe_ms[f] = timed(lambda: one(f), 3)
        results[f] = h.fetch()
        kernels[f] = h.last_kernel()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = h.launch_count()
    barrier()
    # timed region: K steps; an event between launches gives every kernel's duration INSIDE the region
    marks = [[torch.cuda.Event(enable_timing=True) for _ in range(len(fpes) + 1)] for _ in range(args.steps)]
    for k in range(args.steps):
        for i, f in enumerate(fpes):
            marks[k][i].record(stream)
            one(f)
        marks[k][len(fpes)].record(stream)
    marks[-1][-1].synchronize()
    barrier()
    ms_total = max_over_ranks(marks[0][0].elapsed_time(marks[-1][-1]))
    region_ms = {f: max_over_ranks(sum(marks[k][i].elapsed_time(marks[k][i + 1]) for k in range(args.steps)) / args.steps)
                 for i, f in enumerate(fpes)}
    launches = h.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    step_bytes = n_total * bpe * len(fpes)
    value = step_bytes / (ms_step * 1e-3) / 1e9

    value_check, limbs, status = h.fetch()
    sha = limbs_sha(value_check, limbs)
    # all FPE sizes must agree bit for bit, and all ranks with each other
    same = all(results[f][0] == results[fpes[0]][0] and (results[f][1] == results[fpes[0]][1]).all() for f in fpes)
    ranks_identical = None
    if world > 1:
        shas = [None] * world
        dist.all_gather_object(shas, sha)
        ranks_identical = all(s == shas[0] for s in shas)

    # ---- roofline of the dominant kernel (the slowest FPE instantiation), slower of HBM and FP64 ----
    dom = max(fpes, key=lambda f: region_ms[f])
    roofline = roofline_of(args.op, args.dist, dom, ee, n_total, region_ms[dom], peak * world, peak_src,
                           r_fp64 * world, read_gbs * world if read_gbs else None, kernels[dom])
    roofline.update({
        "peak_source": peak_src + (f" x {world} GPUs" if world > 1 else ""),
        "read_stream_GBs": round(read_gbs, 1) if read_gbs else None,
        "per_fpe_GBs": {str(f): round(n_total * bpe / (region_ms[f] * 1e-3) / 1e9, 1) for f in fpes},
        "per_fpe_GBs_burst": {str(f): round(n_total * bpe / (per_fpe_ms[f] * 1e-3) / 1e9, 1) for f in fpes},
        "per_fpe": {str(f): {k: v for k, v in roofline_of(args.op, args.dist, f, ee, n_total, region_ms[f], peak * world, peak_src,
                                                           r_fp64 * world, read_gbs * world if read_gbs else None).items()
                             if k in ("bound", "achieved", "peak", "frac", "fp64_instr_per_elem")} for f in fpes},
        "note": "achieved = algorithmic bytes per launch (n_total * %d B) / mean duration of that kernel's launches inside the "
                "timed region (CUDA events on the launching stream between launches, max over ranks); frac = max(T_hbm, "
                "T_fp64) / T_measured; per_fpe_GBs_burst = the same kernels timed 3 launches at a time before the region "
                "(no power-cap clock sag); N > 1 includes the limb exchange" % bpe})

    # ---- end to end through the synchronous C-ABI call with HOST buffers -----------------------------
    e2e = None
    if not args.no_e2e:
        def e2e_measure(ha, hb, steps):
            def e2e_step():
                out = None
                for f in fpes:
                    if args.op == "exsum":
                        out = h.exsum(n, ha, 1, 0, f, ee)
                    else:
                        out = h.exdot(n, ha, 1, 0, hb, 1, 0, f, ee)
                return out
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                v = e2e_step()
            barrier()
            return max_over_ranks((time.perf_counter() - t0) / steps), v

        e2e_steps = max(1, min(args.steps, 3))
        ha = torch.empty(n, dtype=torch.float64, pin_memory=True)
        ha.copy_(a)
        hb = None
        if args.op == "exdot":
            hb = torch.empty(n, dtype=torch.float64, pin_memory=True)
            hb.copy_(b)
        torch.cuda.synchronize()
        dt, v_e2e = e2e_measure(ha, hb, e2e_steps)
        # pageable host memory: what the reference's callers pass (tests/test.exsum.cpu.cpp:79 _mm_malloc, new[])
        pa = np.empty(n, dtype=np.float64)
        pa[:] = ha.numpy()
        pb = None
        if args.op == "exdot":
            pb = np.empty(n, dtype=np.float64)
            pb[:] = hb.numpy()
        del ha, hb
        dtp, v_pg = e2e_measure(pa, pb, 1 if n_total >= (1 << 29) else e2e_steps)
        del pa, pb
        e2e = {"value": round(step_bytes / dt / 1e9, 2), "unit": "GB/s",
               "h2d_bytes_per_step": n * bpe * len(fpes), "d2h_bytes_per_step": 368 * len(fpes),
               "steps": e2e_steps, "ms_per_step": round(dt * 1e3, 2), "host_memory": "pinned",
               "pageable": {"value": round(step_bytes / dtp / 1e9, 2), "unit": "GB/s", "ms_per_step": round(dtp * 1e3, 2),
                            "matches_device_result": bool(v_pg == results[fpes[-1]][0])},
               "note": "exblas_b200_exsum/exdot with HOST pointers: chunked H2D overlapped with the kernels, result read back "
                       "each call; aggregate over all ranks (every rank streams its own shard of the one vector; with the fused "
                       "transport the closing kernel of each call also exchanges limbs with the peers); h2d / d2h bytes are per rank",
               "matches_device_result": bool(v_e2e == results[fpes[-1]][0])}

    # ---- side measurements, outside the headline region -----------------------------------------------
    extras = None
    if not args.no_extras and args.op == "exsum":
        extras = {}
        rp = dict(peak=peak * world, peak_src=peak_src, r_fp64=r_fp64 * world, read_gbs=read_gbs * world if read_gbs else None)

        def leg(op, dname, nn_total, xa, xb_, variants, rm, known=None, reps=5):
            nn = xa.numel()
            out = {}
            for f, e_ in variants:
                fn = lambda: reduce_async(op, nn, xa, xb_, f, e_, rm)   # noqa: E731
                fn()
                barrier()
                ms = timed(fn, reps)
                v, l, st = h.fetch()
                rec = {"ms": round(ms, 4), "GBs": round(nn_total * (16 if op == "exdot" else 8) / (ms * 1e-3) / 1e9, 1),
                       "value": v, "status": st, "limbs_sha": limbs_sha(v, l),
                       "roofline": {k: val for k, val in roofline_of(op, dname, f, e_, nn_total, ms, kernel=h.last_kernel(), **rp).items()
                                    if k in ("bound", "achieved", "peak", "frac", "fp64_instr_per_elem", "kernel", "frac_of_hbm_copy_peak")}}
                if known is not None:
                    rec["exact"] = bool(v == known and st == 0)
                out[f"fpe{f}{'ee' if e_ else ''}"] = rec
            return out

        # (1) sustained: the headline step repeated for >= 1.2 s of device time (power-cap clocks by construction)
        reps_s = max(3, int(1200.0 / max(ms_step, 1e-3)))
        smp = ClockSampler(local_rank)
        if rank == 0:
            smp.start()
        ms_s = timed(step, reps_s)
        ck = smp.stop() if rank == 0 else None
        extras["sustained"] = {"GBs": round(step_bytes / (ms_s * 1e-3) / 1e9, 1), "seconds": round(ms_s * reps_s * 1e-3, 2),
                               "frac_of_hbm_copy_peak": round(step_bytes / (ms_s * 1e-3) / 1e9 / (peak * world), 4), "clocks": ck}
        del a
        a = None
        torch.cuda.empty_cache()

        # (2) BASELINE config 3: the ill-conditioned ExDOT (cond > 1e32) with a KNOWN exact answer (1.5), the ONE pair of
        #     2^30-element vectors sharded over the ranks
        xa, xb_ = cm.cancelling_pair(n_total, "dot", seed=7, device=dev)
        if world > 1:
            xa, xb_ = xa[lo:hi].clone(), xb_[lo:hi].clone()
            torch.cuda.empty_cache()
        torch.cuda.synchronize()
        extras["exdot_illcond_known_answer"] = leg("exdot", "illcond", n_total, xa, xb_, [(0, False), (3, False), (8, True)],
                                                   xb.ROUND_EXACT, known=1.5)
        extras["exdot_illcond_known_answer"]["workload"] = (
            f"ExDOT of ONE pair of 2^{args.log2n}-element vectors, cancelling ill-conditioned pairs (cond > 1e32), exact result 1.5, "
            f"sharded over {world} GPU(s); aggregate GB/s, 16 B/element")
        del xa, xb_
        torch.cuda.empty_cache()

        if world > 1:
            # (3) weak scaling (round 1's headline): 2^30 elements PER GPU
            wa = gen_sliced(args.dist, n_total * world, rank * n_total, (rank + 1) * n_total, 1, dev)
            torch.cuda.synchronize()

            def wstep():
                for f in fpes:
                    reduce_async("exsum", n_total, wa, None, f, ee, xb.ROUND_REFERENCE)
            wstep()
            barrier()
            msw = timed(wstep, 5)
            extras["weak_scaling"] = {"GBs": round(world * n_total * 8 * len(fpes) / (msw * 1e-3) / 1e9, 1),
                                      "elements_per_gpu": n_total, "ms_per_step": round(msw, 4)}
            # (4) the whole vector on ONE GPU (rank 0) must give the bits of the sharded reduction
            if rank == 0:
                h.set_option("fused_allreduce", 0)
                whole = gen_sliced(args.dist, n_total, 0, n_total, 1, dev, out=wa) if wa.numel() == n_total else None
                h.exsum_async(n_total, whole, 1, 0, fpes[0], ee, xb.ROUND_REFERENCE)
                v1, l1, s1 = h.fetch()
                extras["matches_single_gpu"] = bool(limbs_sha(v1, l1) == sha and s1 == status)
                h.set_option("fused_allreduce", 1 if (red and red.fused) else 0)
            del wa
            torch.cuda.empty_cache()
            # (5) strong scaling in the latency regime: 2^28 and 2^24 elements in total
            mid = {}
            for lg in (24, 28):
                nt = 1 << lg
                l2, h2 = xd.shard_bounds(nt, rank, world)
                ma = gen_sliced("loguniform", nt, l2, h2, 1, dev)
                fn = lambda: reduce_async("exsum", h2 - l2, ma, None, 3, False, xb.ROUND_REFERENCE)   # noqa: E731
                for _ in range(3):
                    fn()
                barrier()
                msm = timed(fn, 50)
                mid[f"2^{lg}"] = {"us": round(msm * 1e3, 2), "GBs": round(nt * 8 / (msm * 1e-3) / 1e9, 1)}
                del ma
            extras["strong_scaling_small"] = mid
        else:
            # (3) FP64-bound and narrow-range variants of ExSUM: naive data (all 1.1), FPE 3 / 4 / 8 without early exit
            #     (every summand walks all levels) and FPE 8 with early exit; the same on the signed log-uniform vector
            na = cm.init_naive(n_total, device=dev)
            extras["exsum_naive"] = leg("exsum", "naive", n_total, na, None, [(0, False), (3, False), (4, False), (8, False), (8, True)],
                                        xb.ROUND_REFERENCE)
            del na
            sa = gen_sliced("loguniform_signed", n_total, 0, n_total, 1, dev)
            extras["exsum_loguniform_signed"] = leg("exsum", "loguniform_signed", n_total, sa, None, [(0, False), (3, False), (8, False)],
                                                    xb.ROUND_REFERENCE)
            del sa
            torch.cuda.empty_cache()
            # (4) BASELINE configs[0] on the GPU: 2^25 doubles, FPE 8 early exit, three distributions; device-resident and
            #     through the host-pointer call (pageable numpy memory)
            c0 = {}
            for dname in ("naive", "loguniform", "illcond"):
                n0 = 1 << 25
                ca = gen(dname, n0, 0, n0, 1, dev)
                fn = lambda: h.exsum_async(n0, ca, 1, 0, 8, True)      # noqa: E731
                for _ in range(3):
                    fn()
                ms0 = timed(fn, 20)
                hc = ca.cpu().numpy()
                h.exsum(n0, hc, 1, 0, 8, True)
                t0 = time.perf_counter()
                for _ in range(3):
                    h.exsum(n0, hc, 1, 0, 8, True)
                th = (time.perf_counter() - t0) / 3
                c0[dname] = {"device_GBs": round(n0 * 8 / (ms0 * 1e-3) / 1e9, 1), "host_pageable_e2e_GBs": round(n0 * 8 / th / 1e9, 2)}
                del ca
            extras["config0_on_gpu"] = {"workload": "ExSUM 2^25 doubles, FPE 8 early-exit (tests/test.exsum.cpu.cpp:107-112)", **c0}
            # (5) BASELINE configs[3], latency regime: device time per reduction in CUDA-graph replay (50 captured calls)
            lat = {}
            la = gen("loguniform", 1 << 24, 0, 1 << 24, 1, dev)
            for lg in (10, 16, 20, 24):
                nn = 1 << lg
                for f, e_, tag in ((0, False, "fpe0"), (3, False, "fpe3")):
                    for _ in range(3):
                        h.exsum_async(nn, la, 1, 0, f, e_)
                    stream.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=stream):
                        for _ in range(50):
                            h.exsum_async(nn, la, 1, 0, f, e_)
                    g.replay()
                    stream.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    for _ in range(10):
                        g.replay()
                    e1.record(stream)
                    e1.synchronize()
                    us = e0.elapsed_time(e1) * 1e3 / 500
                    lat.setdefault(f"2^{lg}", {})[tag] = {"graph_us": round(us, 2), "GBs": round(nn * 8 / us / 1e3, 1)}
                    del g
            extras["latency_graph_replay"] = lat
            del la
            torch.cuda.empty_cache()
            # (6) BASELINE config 5: ExGEMV 32768 x 32768, the reference test's fp-uniform data ("10 0": 10 binades),
            #     'N' and 'T', fpe 0 and 3; GB/s = (m*n + m + n) * 8 / t as the reference reports it (ExGEMV.cpp:208-211)
            if args.log2n >= 30:
                gm = 32768
                A = torch.empty(gm * gm, dtype=torch.float64, device=dev)
                for s in range(0, gm * gm, 1 << 27):
                    A[s:s + (1 << 27)] = cm.init_fpuniform(gm * gm, 10, 5, seed=1, neg_ratio=2, lo=s, hi=s + (1 << 27), device=dev)
                gx = cm.init_fpuniform(gm, 10, 5, seed=2, neg_ratio=2, device=dev)
                gy = torch.zeros(gm, dtype=torch.float64, device=dev)
                gv, ys = {}, {}
                for trans in ("N", "T"):
                    for f in (0, 3):
                        fn = lambda: xb.exgemv(trans, gm, gm, 1.0, A, gm, 0, gx, 1, 0, 0.0, gy, 1, 0, f, False, handle=h, sync=False)   # noqa: E731
                        fn()
                        fn()
                        ms = timed(fn, 5)
                        gb = (gm * gm + 2 * gm) * 8 / (ms * 1e-3) / 1e9
                        gv[f"{trans} fpe{f}"] = {"ms": round(ms, 3), "GBs": round(gb, 1), "frac_of_hbm_copy_peak": round(gb / peak, 4)}
                        ys[(trans, f)] = gy.clone()
                gv["fpe_variants_bit_identical"] = bool((ys[("N", 0)].view(torch.int64) == ys[("N", 3)].view(torch.int64)).all()
                                                        and (ys[("T", 0)].view(torch.int64) == ys[("T", 3)].view(torch.int64)).all())
                extras["exgemv_32768"] = gv
                del A, gx, gy, ys

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            cpu = CpuExsum(args.op)
            lg = cpu_sample_log2n(args, cpu, fpes, 4, 20.0)
            _, _, cpu_baseline = cpu_run(args, fpes, lg, 3, 1, cpu)
        except Exception as ex:  # the checker is optional infrastructure; never fail the bench on it
            cpu_baseline = {"value": None, "unit": "GB/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_of(args, world),
        "shard": {"elements_per_gpu": n, "launches_per_step": launches // max(args.steps, 1),
                  "collective": (args.collective if world > 1 else None)},
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "e2e": e2e,
        "gpu_launches": launches,
        "clocks": clocks,
        "result": {"value": value_check, "value_hex": float(value_check).hex(), "status": status, "limbs_sha": sha,
                   "all_fpe_bit_identical": bool(same), "identical_on_all_ranks": ranks_identical},
        "extras": extras,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    fpes = [int(x) for x in args.fpe.split(",") if x != ""]
    if args.impl == "reference":
        run_reference(args, fpes)
    else:
        run_ours(args, fpes)


if __name__ == "__main__":
    main()
