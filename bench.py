#!/usr/bin/env python
"""bench.py -- ExSUM / ExDOT throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--op exsum|exdot] [--dist loguniform|naive|illcond] [--log2n 30] [--fpe 3,4,8]
    torchrun ... bench.py --gpus N ...          (N > 1: one rank per GPU)

Workload (default = BASELINE.json configs[1]): ExSUM of 2^30 doubles (8 GiB), log-uniform
1e-100..1e100 (the reference's init_fpuniform(n, 664, 332), positive values; --dist
loguniform_signed gives the same magnitudes with random signs), FPE sizes 3, 4 and 8.  One STEP = one
reduction per FPE size over the same resident vector, i.e. 3 kernel launches and 3 x 8 GiB of
algorithmic traffic.  `value` = algorithmic bytes / device time (CUDA events on the launching
stream, inputs resident in HBM, 8 GiB >> 126 MB L2 so every pass streams from DRAM).
`e2e` = the same step through the synchronous C-ABI entry point with PINNED HOST buffers: H2D
copies and the D2H read of the result are inside the timed region.
N > 1: weak scaling -- every rank owns its own 2^30-element shard of an N * 2^30 vector, reduces
it on its GPU and the ranks combine 44 x int64 (limbs + status counters) exactly: by default inside
the closing kernel over NVLink peer memory (--collective fused), or with ncclAllReduce (--collective nccl).

--impl reference times the reference's own CPU ExSUM (oracle/_ref, unmodified sources, OpenMP over
all host cores; falls back to the oracle port if the prebuilt library is absent) on a bounded sample
of the same workload.
"""
from __future__ import annotations

import argparse
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--op", default="exsum", choices=["exsum", "exdot"])
    ap.add_argument("--dist", default="loguniform", choices=["loguniform", "loguniform_signed", "naive", "illcond"])
    ap.add_argument("--log2n", type=int, default=30)
    ap.add_argument("--fpe", default="3,4,8")
    ap.add_argument("--early-exit", type=int, default=0)
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl"],
                    help="N > 1: limb exchange inside the closing kernel over peer memory (fused) or ncclAllReduce")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the ExDOT (BASELINE config 3) side measurement")
    ap.add_argument("--cpu-log2n", type=int, default=27, help="sample size of the CPU baseline / reference arm")
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


def workload_name(args, n):
    d = {"loguniform": "log-uniform 1e-100..1e100 (the reference's init_fpuniform(n,664,332): positive values)",
         "loguniform_signed": "log-uniform 1e-100..1e100 (init_fpuniform(n,664,332) magnitudes, random sign)",
         "naive": "all 1.1 (init_naive)", "illcond": "init_ill_cond(n, 1e32)"}[args.dist]
    return f"{args.op.upper()} n=2^{args.log2n} doubles per GPU, {d}, FPE sizes {args.fpe}" + \
        (" early-exit" if args.early_exit else "")


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
                                          "--format=csv,noheader,nounits", "-lms", "25"],
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
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(op, fpes):
    """dram bytes per launch from the committed ncu capture, if one matches (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    try:
        t = json.load(open(p))
        return t.get(f"{op}_2p30")
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own CPU ExSUM on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_run(args, fpes, log2n, reps):
    """-> (GB/s over the sample, description dict).  Uses oracle/_ref (unmodified reference) when present."""
    import numpy as np
    from oracle.oracle import Oracle, Reference
    n = 1 << log2n
    a = np.ascontiguousarray(gen(args.dist, 1 << args.log2n, 0, n, 1, None))
    b = np.ascontiguousarray(gen(args.dist, 1 << args.log2n, 0, n, 2, None)) if args.op == "exdot" else None
    ee = bool(args.early_exit)
    if args.op == "exsum" and Reference.available():
        ref = Reference()
        kind, cores = "reference", ref.use_all_cores()

        def one(fpe):
            return ref.exsum(a, fpe=fpe, early_exit=ee, parallel=True)
    else:
        O = Oracle()
        kind, cores = "port", (O.use_all_cores() if args.op == "exsum" else 1)

        def one(fpe):
            if args.op == "exsum":
                return O.exsum_parallel(a, fpe=fpe, early_exit=ee)
            return O.exdot(a, b, fpe=fpe, early_exit=ee)[0]
    one(fpes[0])                                        # warm-up
    best = {}
    t_all = []
    for _ in range(reps):
        t0 = time.perf_counter()
        for f in fpes:
            t1 = time.perf_counter()
            one(f)
            best[f] = min(best.get(f, 1e30), time.perf_counter() - t1)
        t_all.append(time.perf_counter() - t0)
    bytes_per_elem = 16 if args.op == "exdot" else 8
    step_bytes = n * bytes_per_elem * len(fpes)
    gbs = step_bytes / min(t_all) / 1e9
    desc = {"value": round(gbs, 3), "unit": "GB/s", "cores": cores, "kind": kind,
            "sample": f"first 2^{log2n} elements of the same vector, FPE {','.join(map(str, fpes))}"
                      f"{' early-exit' if ee else ''}, min of {reps} passes; per-FPE GB/s: " +
                      ", ".join(f"{f}:{n * bytes_per_elem / best[f] / 1e9:.2f}" for f in fpes)}
    return gbs, min(t_all), desc


def run_reference(args, fpes):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    reps = max(1, min(args.steps, 5))
    gbs, t_step, desc = cpu_run(args, fpes, args.cpu_log2n, reps)
    line = {
        "impl": "reference", "metric": "ExSUM/ExDOT GB/s", "value": round(gbs, 3), "unit": "GB/s",
        "n_gpus": args.gpus, "steps": reps, "warmup": 1, "ms_per_step": round(t_step * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, 1 << args.log2n),
                   "note": "reference CPU implementation on the host cores; each step is a bounded sample "
                           f"(2^{args.cpu_log2n} elements) of the workload"},
        "cpu_baseline": desc,
        "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, fpes):
    import torch
    import exblas_b200 as xb
    from exblas_b200 import dist as xd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    ee = bool(args.early_exit)
    n = 1 << args.log2n
    n_total = n * world
    bytes_per_elem = 16 if args.op == "exdot" else 8

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

    a = gen(args.dist, n_total, rank * n, (rank + 1) * n, 1, dev)
    b = gen(args.dist, n_total, rank * n, (rank + 1) * n, 2, dev) if args.op == "exdot" else None
    torch.cuda.synchronize()

    def one(fpe):
        if args.op == "exsum":
            h.exsum_async(n, a, 1, 0, fpe, ee, xb.ROUND_REFERENCE)
        else:
            h.exdot_async(n, a, 1, 0, b, 1, 0, fpe, ee, xb.ROUND_REFERENCE)
        if world > 1:
            h.allreduce_async(xb.ROUND_REFERENCE)

    def step():
        for f in fpes:
            one(f)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    results = {}
    # per-variant timing (outside the headline region): average kernel duration per FPE size
    per_fpe_ms = {}
    for f in fpes:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            one(f)
        e1.record(stream)
        e1.synchronize()
        per_fpe_ms[f] = e0.elapsed_time(e1) / 3
        results[f] = h.fetch()
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
    ms_total = marks[0][0].elapsed_time(marks[-1][-1])
    region_ms = {f: sum(marks[k][i].elapsed_time(marks[k][i + 1]) for k in range(args.steps)) / args.steps
                 for i, f in enumerate(fpes)}
    launches = h.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    step_bytes = n_total * bytes_per_elem * len(fpes)
    value = step_bytes / (ms_step * 1e-3) / 1e9

    value_check, limbs, status = h.fetch()
    # all FPE sizes must agree bit for bit (and across ranks when N > 1)
    same = all(results[f][0] == results[fpes[0]][0] and (results[f][1] == results[fpes[0]][1]).all() for f in fpes)

    # ---- roofline of the dominant kernel (the slowest FPE instantiation) ------------------------
    peak, peak_src = measured_peak()
    dom = max(fpes, key=lambda f: region_ms[f])
    launch_bytes = n * bytes_per_elem
    achieved = launch_bytes / (region_ms[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": ncu_traffic(args.op, fpes),
                "kernel": f"exblas_reduce_kernel<F={dom},EE={int(ee)},DOT={int(args.op == 'exdot')}>",
                "peak_source": peak_src,
                "per_fpe_GBs": {str(f): round(launch_bytes / (region_ms[f] * 1e-3) / 1e9, 1) for f in fpes},
                "per_fpe_GBs_burst": {str(f): round(launch_bytes / (per_fpe_ms[f] * 1e-3) / 1e9, 1) for f in fpes},
                "note": "achieved = algorithmic bytes per launch (n * %d B) / mean duration of that kernel's launches inside "
                        "the timed region (CUDA events on the launching stream between launches); per_fpe_GBs_burst = "
                        "the same kernels timed 3 launches at a time before the region (no power-cap clock sag); includes "
                        "the NCCL limb all-reduce when N > 1" % bytes_per_elem}

    # ---- end to end through the synchronous C-ABI call with pinned host buffers ----------------
    e2e = None
    if not args.no_e2e:
        ha = torch.empty(n, dtype=torch.float64, pin_memory=True)
        ha.copy_(a)
        hb = None
        if args.op == "exdot":
            hb = torch.empty(n, dtype=torch.float64, pin_memory=True)
            hb.copy_(b)
        torch.cuda.synchronize()

        def e2e_step():
            out = None
            for f in fpes:
                if args.op == "exsum":
                    out = h.exsum(n, ha, 1, 0, f, ee)
                else:
                    out = h.exdot(n, ha, 1, 0, hb, 1, 0, f, ee)
            return out

        e2e_steps = max(1, min(args.steps, 3))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            v_e2e = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": round(step_bytes / dt / 1e9, 2), "unit": "GB/s",
               "h2d_bytes_per_step": n * bytes_per_elem * len(fpes), "d2h_bytes_per_step": 368 * len(fpes),
               "steps": e2e_steps, "ms_per_step": round(dt * 1e3, 2),
               "note": "exblas_b200_exsum/exdot (host pointers, pinned): chunked H2D overlapped with the kernels, "
                       "result read back each call; N > 1: every rank streams its own host shard, and with the fused transport the "
                       "closing kernel of each call also exchanges limbs with the peers",
               "matches_device_result": bool(v_e2e == results[fpes[-1]][0]) if world == 1 else None}
        del ha, hb

    # ---- side measurement, outside the headline region: BASELINE config 3, the ill-conditioned
    # ExDOT (cond > 1e32) with a KNOWN exact answer: every rank's shard dots to exactly 1.5 ----------
    extras = None
    if not args.no_extras and args.op == "exsum":
        del a
        torch.cuda.empty_cache()
        from exblas_b200 import common as cm
        xa, xb_ = cm.cancelling_pair(n, "dot", seed=7 + rank, device=dev)
        torch.cuda.synchronize()
        ex = {}
        for f, e_ in ((0, False), (3, False), (8, True)):
            def dot_once():
                h.exdot_async(n, xa, 1, 0, xb_, 1, 0, f, e_, xb.ROUND_EXACT)
                if world > 1:
                    h.allreduce_async(xb.ROUND_EXACT)
            dot_once()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            d0.record(stream)
            for _ in range(5):
                dot_once()
            d1.record(stream)
            d1.synchronize()
            ms = d0.elapsed_time(d1) / 5
            v, _, st = h.fetch()
            ex[f"fpe{f}{'ee' if e_ else ''}"] = {"GBs": round(n_total * 16 / (ms * 1e-3) / 1e9, 1), "value": v,
                                                 "exact": bool(v == 1.5 * world and st == 0)}
        extras = {"exdot_illcond_known_answer": ex,
                  "workload": f"ExDOT n=2^{args.log2n} per GPU, cancelling ill-conditioned pairs (cond > 1e32), "
                              f"exact result {1.5 * world}; aggregate GB/s over {world} GPU(s), 16 B/element"}
        del xa, xb_
        torch.cuda.empty_cache()
        # BASELINE config 5: ExGEMV 32768 x 32768, the reference test's fp-uniform data ("10 0": 10 binades),
        # 'N' and 'T', superaccumulator-only (register window) and FPE 3; GB/s = (m*n + m + n) * 8 / t as the
        # reference reports it (ExGEMV.cpp:208-211).  One GPU only (the matrix is not sharded).
        if world == 1 and args.log2n >= 30:
            gm = 32768
            A = torch.empty(gm * gm, dtype=torch.float64, device=dev)
            for lo in range(0, gm * gm, 1 << 27):
                A[lo:lo + (1 << 27)] = cm.init_fpuniform(gm * gm, 10, 5, seed=1, neg_ratio=2, lo=lo, hi=lo + (1 << 27), device=dev)
            gx = cm.init_fpuniform(gm, 10, 5, seed=2, neg_ratio=2, device=dev)
            gy = torch.zeros(gm, dtype=torch.float64, device=dev)
            gv = {}
            ys = {}
            for trans in ("N", "T"):
                for f in (0, 3):
                    for _ in range(2):
                        xb.exgemv(trans, gm, gm, 1.0, A, gm, 0, gx, 1, 0, 0.0, gy, 1, 0, f, False, handle=h, sync=False)
                    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    d0.record(stream)
                    for _ in range(5):
                        xb.exgemv(trans, gm, gm, 1.0, A, gm, 0, gx, 1, 0, 0.0, gy, 1, 0, f, False, handle=h, sync=False)
                    d1.record(stream)
                    d1.synchronize()
                    ms = d0.elapsed_time(d1) / 5
                    gv[f"{trans} fpe{f}"] = {"ms": round(ms, 3), "GBs": round((gm * gm + 2 * gm) * 8 / (ms * 1e-3) / 1e9, 1)}
                    ys[(trans, f)] = gy.clone()
            gv["fpe_variants_bit_identical"] = bool((ys[("N", 0)].view(torch.int64) == ys[("N", 3)].view(torch.int64)).all()
                                                    and (ys[("T", 0)].view(torch.int64) == ys[("T", 3)].view(torch.int64)).all())
            extras["exgemv_32768"] = gv
            del A, gx, gy, ys

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            _, _, cpu_baseline = cpu_run(args, fpes, args.cpu_log2n, 3)
        except Exception as ex:  # the checker is optional infrastructure; never fail the bench on it
            cpu_baseline = {"value": None, "unit": "GB/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}

    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": "ExSUM/ExDOT GB/s", "value": round(value, 1), "unit": "GB/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, n), "elements_per_gpu": n, "launches_per_step": launches // args.steps,
                   "l2": "inputs (8 GiB per pass) are far larger than the 126 MB L2; no flush needed",
                   "parallelism": f"shard{world}" if world > 1 else "single",
                   "collective": (args.collective if world > 1 else None)},
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "e2e": e2e,
        "gpu_launches": launches,
        "clocks": clocks,
        "result": {"value": value_check, "status": status, "all_fpe_bit_identical": bool(same)},
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
