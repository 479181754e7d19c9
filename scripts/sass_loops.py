"""Development helper: per-loop instruction statistics of one kernel's SASS (cuobjdump -sass), to see whether
local-memory traffic (register spills: LDL / STL) sits inside a hot loop.

    cuobjdump -sass lib.so | python scripts/sass_loops.py <substring of the mangled kernel name>
"""
import re, sys
name = sys.argv[1]
lines = sys.stdin.read().splitlines()
ins = []
on = False
for ln in lines:
    if "Function :" in ln:
        on = name in ln
        continue
    if not on:
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.\S+)?\s+(?:\S+,\s*)?(0x[0-9a-f]+)", t)
    if m and not t.startswith("BRA.DIV"):
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr:
            loops.append((addr[tgt], i))
print(f"{len(ins)} instructions, {len(loops)} backward branches")
for lo, hi in sorted(loops):
    body = [t for _, t in ins[lo:hi + 1]]
    cnt = lambda pat: sum(1 for t in body if re.search(pat, t))
    print(f"loop {ins[lo][0]:#07x}..{ins[hi][0]:#07x}: {len(body):5d} instr  DADD/DMUL/DFMA {cnt(r'^(@!?U?P\d+\s+)?D(ADD|MUL|FMA)'):4d}  LDG {cnt(r'LDG'):3d}  LDS {cnt(r'LDS'):3d}  STS {cnt(r'STS'):3d}  "
          f"LDL {cnt(r'LDL'):3d}  STL {cnt(r'STL'):3d}  CALL {cnt(r'CALL'):2d}  BAR {cnt(r'BAR'):2d}")
