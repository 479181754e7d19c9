"""exblas_b200/csrc/window.cuh on the host: the register window of the superaccumulator (digit split by
magic-constant adds, drain, re-anchoring policy) is plain IEEE arithmetic, so the very header the
CUDA kernels include is compiled with g++ and checked against the ordinary limb accumulation on
~850 streams (narrow / wide / moving ranges, zeros, outliers, window edges, both signs)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_window_arithmetic_on_host(tmp_path):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    exe = str(tmp_path / "window_host_check")
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-o", exe,
                           os.path.join(HERE, "window_host_check.cpp")])
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout
