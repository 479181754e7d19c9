"""SURVEY section 8f rank 4 -- the reference's OWN GPU test mains (tests/test.exsum.gpu.cpp,
test.exdot.gpu.cpp, test.exgemv.gpu.cpp), compiled UNMODIFIED against include/blas1.hpp / blas2.hpp
of this repository and linked to libexblas_b200.so by `make -C oracle ref_tests` (in the build
container, where /root/reference exists; the binaries travel in oracle/_ref/).  Run with the
arguments of the reference's CTest files (src/gpu/blas/blas1/CMakeLists.txt:9-29,
src/gpu/blas/blas2/CMakeLists.txt:12-63); pass == the reference's own "TestPassed; ALL OK!".

Two builds: the default one checks every FPE variant against the superaccumulator-only result;
the -DEXBLAS_VS_MPFR one checks against MPFR within 1e-16 and is run with EXBLAS_B200_ROUND=exact
(with the reference's own Round() that mode can miss by 1 ulp)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")

BLAS1_ARGS = [["22"], ["22", "2", "0", "n"], ["22", "50", "0", "n"], ["22", "1e+50", "0", "i"], ["22", "100", "50"]]
GEMV_ARGS = [[t, m, n] + rest for t in ("N", "T") for (m, n) in (("512", "512"), ("512", "1024"), ("1024", "512"))
             for rest in ([], ["50", "0", "n"], ["10", "0", "y"], ["1e+50", "0", "i"])]


def run(exe, args, env_extra=None):
    path = os.path.join(REF, exe)
    if not os.path.exists(path):
        pytest.skip(f"{exe} not prebuilt (make -C oracle ref_tests needs /root/reference)")
    env = dict(os.environ)
    env.update(env_extra or {})
    p = subprocess.run([path] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=600)
    assert p.returncode == 0, (exe, args, p.stdout[-2000:], p.stderr[-2000:])
    assert "TestPassed; ALL OK!" in p.stdout, (exe, args, p.stdout[-2000:])
    return p.stdout


@pytest.mark.parametrize("exe", ["test.exsum.gpu", "test.exdot.gpu"])
def test_reference_blas1_mains_self_consistency(gpu, exe):
    for args in BLAS1_ARGS:
        run(exe, args)


@pytest.mark.parametrize("exe", ["test.exsum.gpu.mpfr", "test.exdot.gpu.mpfr"])
def test_reference_blas1_mains_vs_mpfr(gpu, exe):
    for args in BLAS1_ARGS:
        a = list(args)
        a[0] = "20"                                  # the MPFR loop is serial
        run(exe, a, {"EXBLAS_B200_ROUND": "exact"})


def test_reference_exgemv_main(gpu):
    for args in GEMV_ARGS:
        run("test.exgemv.gpu", args)
    for args in GEMV_ARGS[::3]:
        run("test.exgemv.gpu.mpfr", args, {"EXBLAS_B200_ROUND": "exact"})
