import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle, build
    build()
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference CPU library, when its prebuilt .so (or /root/reference) is present."""
    from oracle.oracle import Reference, build
    build()
    if not Reference.available():
        pytest.skip("oracle/_ref/libexblas_ref.so not present (no /root/reference on this machine)")
    return Reference()


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "golden.npz")
    return np.load(path)


@pytest.fixture(scope="session")
def lib():
    """The product C-ABI library; built if stale.  No fallback: a build failure fails the tests."""
    from exblas_b200 import build as b
    b.build()
    from exblas_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def gpu():
    """A handle on cuda:0 plus torch for device buffers."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from exblas_b200 import build as b
    b.build()
    import exblas_b200 as xb
    h = xb.Handle(0)
    # By default ExSUM vectors of up to 2^20 elements all take ONE kernel whatever fpe says (option track_max_elems);
    # the parity tests switch that off so that their (small) inputs reach the kernel each fpe / early-exit value selects.
    # test_small_and_mid_sizes_every_shape_regime covers the default routing.
    h.set_option("track_max_elems", 0)
    return h
