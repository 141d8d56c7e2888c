import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a fresh checkout has no built artefacts (they are git-ignored): build them once, as the driver's build() does
    if not os.path.exists(os.path.join(ROOT, "rl-rust_b200", "librlb.so")) or not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def rlb():
    """The product package.  Importing it requires the built librlb.so (no fallback)."""
    import importlib
    return importlib.import_module("rl-rust_b200")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle_py
    oracle_py.lib()
    return oracle_py
