"""Development helper: print per-instruction samples/stalls from `ncu --page source --csv` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 10**12
hdr = None
n = 0
base = None
for r in rows:
    if r and r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}
        n += 1
        if n > 1: break
        continue
    if hdr is None or len(r) <= hdr["stall_wait"]: continue
    try:
        s = int(r[hdr["# Samples"]] or 0)
    except ValueError:
        continue
    try:
        k = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
    except ValueError:
        continue
    if base is None: base = k
    k -= base
    if not (lo <= k <= hi): continue
    print(f"{k:6x} {s:6d} {r[hdr['Instructions Executed']]:>9s} lsb={r[hdr['stall_long_sb']]:>5s} wait={r[hdr['stall_wait']]:>5s} ssb={r[hdr['stall_short_sb']]:>4s} br={r[hdr['stall_branch_resolving']]:>4s}  {r[1][:100]}")
