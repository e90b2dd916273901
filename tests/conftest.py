import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracles():
    """Build (if needed) and load the CPU oracles.  'port' always; 'strict'/'fast' where available."""
    from oracle import pyoracle as po
    po.build(ref=True, port=True)
    out = {"port": po.load("port")}
    for kind in ("strict", "fast"):
        if po.available(kind):
            out[kind] = po.load(kind)
    return out


@pytest.fixture(scope="session")
def port(oracles):
    return oracles["port"]


@pytest.fixture(scope="session")
def engine_lib():
    """The CUDA shared library; built in-tree if stale (nvcc cross-compiles without a GPU)."""
    from spandsp_b200 import build as b
    b.build()
    from spandsp_b200 import engine
    return engine


@pytest.fixture(scope="session")
def gpu_ctx(engine_lib):
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    torch.cuda.init()
    ctx = engine_lib.Context(0)
    yield ctx
    ctx.close()
