"""Golden outputs for the reference's RNGExample.

Runs the reference's UNMODIFIED CPU RNGExample (src/cpu/examples/RNGExample/StrongReproducibility/
RNGExample.cpp) linked to the unmodified reference CPU library (oracle/_ref/RNGExample.cpu, built by
`make -C oracle ref` in the build container) and stores, per parameter set, the lines
"ExBLAS reproducible sum (<variant>): ..." it prints.  The GPU twin of the same example
(src/gpu/examples/RNGExample/..., same mt19937 generator and seed => same elements) built against
libexblas_b200.so must print the same lines (tests/test_reference_mains.py).

    python tests/golden/make_golden_rng.py        # writes tests/golden/rngexample.json
"""
import json
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EXE = os.path.join(ROOT, "oracle", "_ref", "RNGExample.cpu")

CASES = [
    [],                                                          # the CTest invocation: defaults (n = 1000, r = 100)
    ["-n", "100000", "-r", "20", "-l", "-30", "-h", "30"],
    ["-n", "1000003", "-r", "5", "-l", "-126", "-h", "127", "-s", "7"],
    ["-n", "4099", "-r", "50", "-l", "0", "-h", "0", "-s", "3"],
]

LINE = re.compile(r"^ExBLAS reproducible sum \((.*?)\): (.*)$")


def exblas_lines(text):
    return [m.group(0) for m in map(LINE.match, text.splitlines()) if m]


def main():
    out = []
    for args in CASES:
        p = subprocess.run([EXE] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, check=True)
        lines = exblas_lines(p.stdout)
        assert len(lines) == 4, p.stdout
        assert not re.search(r"ExBLAS reproducible sum .* not reproducible", p.stdout)
        out.append({"args": args, "lines": lines})
    with open(os.path.join(HERE, "rngexample.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote {len(out)} cases")


if __name__ == "__main__":
    main()
