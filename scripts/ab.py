"""Development helper: A/B two library builds on the same box (alternating, several rounds)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = {"A(default)": None, "B": sys.argv[1]}
for rnd in range(2):
    for name, lib in libs.items():
        env = dict(os.environ)
        if lib: env["EXBLAS_B200_LIB"] = os.path.join(ROOT, lib)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sweep_opts.py"), "30"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout.strip().splitlines()
        print(name, out[0] if out else "no output", flush=True)
