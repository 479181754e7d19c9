"""Parse `nvcc -Xptxas -v` output into {demangled kernel/function: (registers, stack, spill stores, spill loads)}
and print a table, or the differences between two builds:   python scripts/ptxas_table.py new.txt [old.txt]"""
import re, subprocess, sys

def parse(path):
    out, name = {}, None
    txt = open(path).read().splitlines()
    names = []
    for ln in txt:
        m = re.search(r"Compiling entry function '(\S+)'|Function properties for (\S+)", ln)
        if m:
            name = m.group(1) or m.group(2)
            out.setdefault(name, {})
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
        if m and name:
            out[name].update(stack=int(m.group(1)), sst=int(m.group(2)), sld=int(m.group(3)))
        m = re.search(r"Used (\d+) registers", ln)
        if m and name:
            out[name]["regs"] = int(m.group(1))
    dem = subprocess.run(["c++filt"], input="\n".join(out), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    return {d: v for d, v in zip(dem, out.values())}

def short(n):
    n = re.sub(r"exb::", "", n)
    return re.sub(r"\(.*", "", n)[:90]

new = parse(sys.argv[1])
old = parse(sys.argv[2]) if len(sys.argv) > 2 else None
for k, v in sorted(new.items()):
    if "regs" not in v and not v.get("sst") and not v.get("stack"):
        continue
    line = f"{short(k):92s} regs={v.get('regs','-'):>4} stack={v.get('stack',0):>4} spill={v.get('sst',0)}/{v.get('sld',0)}"
    if old is not None:
        o = old.get(k)
        if o is None:
            line += "   (new)"
        elif (o.get('regs'), o.get('sst', 0), o.get('stack', 0)) != (v.get('regs'), v.get('sst', 0), v.get('stack', 0)):
            line += f"   was regs={o.get('regs','-')} stack={o.get('stack',0)} spill={o.get('sst',0)}/{o.get('sld',0)}"
        else:
            continue
    print(line)
