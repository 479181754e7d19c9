"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, its host-only limb functions agree with the oracle, and -- without a GPU -- the compute
entry points fail loudly instead of falling back to anything."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import same_double
from exblas_b200 import common as cm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "exblas_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(exblas_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/exblas_b200.h but not exported"


def test_binding_table_matches_header(lib):
    from exblas_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_library_does_not_link_oracle(lib):
    """the product must not depend on anything under oracle/"""
    from exblas_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "oracle" not in out and "exblas_ref" not in out
    syms = subprocess.run(["nm", "-D", "--undefined-only", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "oracle_" not in syms and "ref_exsum" not in syms
    for root, _, files in os.walk(os.path.join(ROOT, "exblas_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_host_round_matches_oracle(lib, oracle, golden):
    import exblas_b200 as xb
    for name in golden["sum_cases"]:
        a = golden[f"{name}/a"]
        v0, l = oracle.exsum(a, fpe=0, round_mode=0)
        v1, _ = oracle.exsum(a, fpe=0, round_mode=1)
        assert same_double(xb.round_limbs(l, xb.ROUND_REFERENCE), v0), name
        assert same_double(xb.round_limbs(l, xb.ROUND_REFERENCE), float(golden[f"{name}/ref_round"][0])), name
        assert same_double(xb.round_limbs(l, xb.ROUND_EXACT), v1), name
        assert same_double(xb.round_limbs(l, xb.ROUND_EXACT), float(golden[f"{name}/mpfr"][0])), name


def test_host_normalize_and_merge(lib, oracle):
    import exblas_b200 as xb
    a = cm.init_fpuniform(3000, 664, 332, seed=21, neg_ratio=2)
    whole_v, whole_l = oracle.exsum(a, fpe=0)
    parts = np.array_split(a, 5)
    acc = np.zeros(39, dtype=np.int64)
    for p in parts:
        acc = xb.merge_limbs(acc, oracle.exsum(p, fpe=0)[1])
    assert (acc == whole_l).all()
    # un-normalised but equal value -> same normal form
    l2 = whole_l.copy()
    l2[10] += 7 << 52
    l2[11] -= 7
    n2, neg = xb.normalize_limbs(l2)
    assert (n2 == whole_l).all() and neg == (whole_l[-1] < 0)
    assert same_double(xb.round_limbs(l2, 0), whole_v)


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import exblas_b200 as xb
    h = C.c_void_p()
    rc = lib.exblas_b200_create(C.byref(h), 0)
    assert rc == -3 and not h.value                     # EXBLAS_B200_ENOGPU
    with pytest.raises(xb.ExblasB200Error):
        xb.exsum(8, np.ones(8), 1, 0, 4)
    with pytest.raises(xb.ExblasB200Error):
        xb.exdot(8, np.ones(8), 1, 0, np.ones(8), 1, 0, 4)


def test_reference_error_behaviour(lib):
    """fpe < 0: message on stderr + exit(1) (cpu ExSUM.cpp:25-28); exdot Ng <= 0 -> 0.0 (ExDOT.cpp:70-71)."""
    import exblas_b200 as xb
    with pytest.raises(SystemExit) as e:
        xb.exsum(8, np.ones(8), 1, 0, -1)
    assert e.value.code == 1
    assert xb.exdot(0, np.ones(1), 1, 0, np.ones(1), 1, 0, 4) == 0.0
    assert xb.exdot(-5, np.ones(1), 1, 0, np.ones(1), 1, 0, 4) == 0.0


def test_blas1_hpp_is_source_compatible():
    """A translation unit written against the reference's blas1.hpp compiles against ours unchanged."""
    src = r'''
    #include "blas1.hpp"
    double f(int n, double* a, double* b) {
        double s = exsum(n, a, 1, 0, 8, true);
        s += exsum(n, a, 1, 0, 4);
        s += exsum(n, a, 1, 0, 0, false, false);
        s += exdot(n, a, 1, 0, b, 1, 0, 3);
        s += exdot(n, a, 1, 0, b, 1, 0, 8, true);
        return s;
    }'''
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c++", "-"],
                       input=src, text=True, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr
