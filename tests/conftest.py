"""pytest configuration: registers the `gpu` marker (tests that need a real B200) and skips
those tests automatically when no CUDA device is present, so `-m "not gpu"` and a plain run both
work in the CPU-only build container."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Build (incrementally) the core library, torch shims and the oracle once per session.
    GSR_TEST_OPTS="key=value,..." applies library options (gsr_set_option) for the whole session, so
    the parity suite can be run against a non-default kernel variant."""
    import __graft_entry__ as ge
    ge.build_product()
    opts = os.environ.get("GSR_TEST_OPTS", "")
    if opts:
        import ctypes
        lib = ctypes.CDLL(ge.core_library_path())
        for kv in opts.split(","):
            k, v = kv.split("=")
            assert lib.gsr_set_option(k.encode(), int(v)) >= 0, "unknown option " + k
    return ge
