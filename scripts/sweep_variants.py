"""Development helper: time a set of configurations for each tuning build in build/lib_T*_U*.so."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFGS = [("exsum", "logu", 0, 0), ("exsum", "naive", 2, 0), ("exsum", "naive", 3, 0), ("exsum", "naive", 4, 0),
        ("exsum", "naive", 8, 0), ("exsum", "naive", 8, 1), ("exdot", "logu", 0, 0), ("exdot", "ill", 3, 0), ("exdot", "ill", 8, 1)]
libs = [None] + sorted(glob.glob(os.path.join(ROOT, "build", "lib_T*_U*.so")))
for lib in libs:
    env = dict(os.environ)
    name = "default(T512,U4)"
    T = 512
    if lib:
        env["EXBLAS_B200_LIB"] = lib
        name = os.path.basename(lib)
        T = int(name.split("_T")[1].split("_")[0])
    for op, kind, fpe, ee in CFGS:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "prof_one.py"), op, kind, str(fpe), str(ee), "28", str(T), "5"],
                             env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout.strip().splitlines()
        print(f"{name:18s} {out[-1] if out else 'no output'}", flush=True)
