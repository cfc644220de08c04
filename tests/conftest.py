import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a machine without a CUDA device; on a GPU box nothing is skipped,
    and a missing library is an error there, never a skip."""
    try:
        from scasa_b200 import lib
        have = lib.load().scasa_device_count() > 0
    except Exception:
        have = os.path.exists("/dev/nvidia0")          # the library must load on a GPU box: let the tests fail loudly
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device (run on the B200 box with -m gpu)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def xm():
    from scasa_b200.xmatrix import XMatrix
    return XMatrix.load_npz(os.path.join(GOLDEN, "xmatrix_hg38_alevin.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def c1(xm):
    """Config C1: 200 synthetic cells on the shipped X-matrix, seed 1 (level-Y counts as CSR)."""
    from scasa_b200 import synth
    gen = synth.CellGenerator(xm, seed=1)
    return gen.counts_csr(0, 200)
