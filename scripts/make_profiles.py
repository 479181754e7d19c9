"""Turn ncu captures (.ncu-rep, tens of MB each) into the small text summaries that are committed under profiles/.

    python scripts/make_profiles.py r02 [dir with the .ncu-rep files = gpurun_out] [output dir = profiles]

(scripts/ncu_round.sh runs it ON the GPU box, so that only the summaries travel back.)

For each gpurun_out/prof_<round>_*.ncu-rep: profiles/<name>.metrics.txt (key raw metrics) and
profiles/<name>.hot_sass.txt (the 40 instructions with the most stall samples).  Also copies the
launch list (exblas kernels + a one-line share summary) and writes profiles/traffic.json (DRAM
bytes per launch, read + write, for bench.py's roofline.traffic)."""
import csv, glob, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
in_dir = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out")
out_dir = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
        "sm__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__pcsamp_sample_count"]

traffic = {}
for rep in sorted(glob.glob(os.path.join(in_dir, f"prof_{rnd}_*.ncu-rep"))):
    name = os.path.basename(rep)[:-len(".ncu-rep")]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {name}: ncu --set full --clock-control none, one launch", f"kernel: {r[ix['Kernel Name']]}"]
    for k in KEYS:
        if k in ix:
            lines.append(f"{k:90s} {r[ix[k]]:>16s} {units[ix[k]]}")
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
            try:
                v = float(r[ix[h]])
            except ValueError:
                continue
            if v >= 0.1:
                lines.append(f"{'stall/issue: ' + h[34:-23]:90s} {v:16.3f}")
    def num(k):
        v, u = float(r[ix[k]].replace(",", "")), units[ix[k]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    tr = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    lines.append(f"{'dram read + write bytes per launch':90s} {tr:16.0f} byte")
    key = name.split(f"{rnd}_")[1]
    traffic[key] = tr
    open(os.path.join(out_dir, name + ".metrics.txt"), "w").write("\n".join(lines) + "\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    hd = None
    body = []
    for row in srows:
        if row and row[0] == "Address":
            hd = {h: i for i, h in enumerate(row)}
            continue
        if hd and len(row) > hd["stall_wait"]:
            try:
                body.append((int(row[hd["# Samples"]] or 0), row))
            except ValueError:
                pass
    tot = sum(s for s, _ in body) or 1
    body.sort(key=lambda t: -t[0])
    cols = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_math", "stall_mio", "stall_branch_resolving", "stall_no_inst"]
    hl = [f"# {name}: instructions with the most warp-stall samples (total {tot})",
          "samples  share  " + " ".join(c[6:] for c in cols) + "  SASS"]
    for s, row in body[:40]:
        hl.append(f"{s:7d} {100 * s / tot:5.1f}%  " + " ".join(f"{row[hd[c]]:>6s}" for c in cols) + "  " + row[1].strip())
    open(os.path.join(out_dir, name + ".hot_sass.txt"), "w").write("\n".join(hl) + "\n")
    print("wrote", name)

# bench.py looks the traffic of a variant up by "op|dist|fpe|ee|log2n" (roofline.traffic); the capture names are
# <op>_fpe<F>[ee]_<dist>_2p<log2n>
import re
old_tj = {}
tpath = os.path.join(ROOT, "profiles", "traffic.json")
if os.path.exists(tpath):          # captures of earlier sessions whose .ncu-rep files are gone: keep their entries
    try:
        old_tj = json.load(open(tpath))
    except Exception:
        old_tj = {}
tj = {k: v for k, v in old_tj.items() if "|" in k}
allc = {**old_tj.get("all", {}), **traffic}
for k, v in traffic.items():
    m = re.match(r"(exsum|exdot)_fpe(\d+)(ee)?_(.+)_2p(\d+)$", k)
    if m:
        dist = {"cancel": "illcond"}.get(m.group(4), m.group(4))
        tj[f"{m.group(1)}|{dist}|{m.group(2)}|{1 if m.group(3) else 0}|{m.group(5)}"] = v
tj["all"] = allc
json.dump(tj, open(os.path.join(out_dir, "traffic.json"), "w"), indent=1)

lc = os.path.join(in_dir, f"launches_{rnd}.csv")
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 5]
    hdr = rows[0]
    i_name, i_val, i_unit, i_id = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    sel = [r for r in rows[1:] if "exblas" in r[i_name] or "exb::" in r[i_name]]
    other = [r for r in rows[1:] if r not in sel]
    def us(r):
        v = float(r[i_val].replace(",", ""))
        return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(r[i_unit], 1)
    with open(os.path.join(out_dir, f"launches_{rnd}.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline\n")
        f.write("# exblas kernels only (the other launches are torch kernels generating the synthetic input, outside the timed region)\n")
        f.write(f"# exblas launches: {len(sel)}, total {sum(map(us, sel)):.1f} us; torch set-up launches: {len(other)}, total {sum(map(us, other)):.1f} us\n")
        f.write("# timed region = the last 6 exblas_reduce_kernel launches (2 steps x FPE 3, 4, 8); before them: warm-up steps, the per-FPE burst timing and the two microbenchmarks\n")
        f.write("id,kernel,duration_us\n")
        for r in sel:
            f.write(f"{r[i_id]},\"{r[i_name]}\",{us(r):.2f}\n")
    print("wrote launches")
