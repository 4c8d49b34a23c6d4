import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_pkg():
    """The product package lives in `scroll-prover_b200/` (hyphenated, as the layout requires)."""
    import importlib

    name = "scroll-prover_b200"
    if name in sys.modules:
        return sys.modules[name]
    mod = importlib.import_module(name)
    sys.modules.setdefault("scroll_prover_b200", mod)
    return mod


@pytest.fixture(scope="session")
def zk():
    return load_pkg()


@pytest.fixture(scope="session")
def ctx(zk):
    c = zk.Context(0)
    yield c
    c.close()
