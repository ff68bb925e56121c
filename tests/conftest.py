import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

# the product must never exit() under pytest; device errors come back as NULL / exceptions
os.environ.setdefault("SMJ_NO_EXIT", "1")
os.environ.setdefault("SMJ_QUIET", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def smj():
    """The product: ctypes binding of libsmj_b200.so (built in-tree by __graft_entry__.build())."""
    so = os.path.join(ROOT, "avx-sort-merge-joins_b200", "libsmj_b200.so")
    if not os.path.exists(so):
        import __graft_entry__
        __graft_entry__.build()
    return importlib.import_module("avx-sort-merge-joins_b200")


@pytest.fixture(scope="session")
def orc():
    import oracle_lib
    return oracle_lib.oracle()


@pytest.fixture(scope="session")
def ref():
    import oracle_lib
    r = oracle_lib.reference()
    if r is None:
        pytest.skip("oracle/_ref/libsmj_ref.so not built (needs /root/reference: make -C oracle ref)")
    return r


@pytest.fixture(scope="session")
def refmat():
    import oracle_lib
    r = oracle_lib.reference_mat()
    if r is None:
        pytest.skip("oracle/_ref/libsmj_ref_mat.so not built (needs /root/reference: make -C oracle ref)")
    return r


@pytest.fixture(scope="session")
def refk8():
    import oracle_lib
    r = oracle_lib.reference_key8b()
    if r is None:
        pytest.skip("oracle/_ref/libsmj_ref_key8b.so not built (needs /root/reference: make -C oracle ref)")
    return r
