import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _host_api():
    import complogic_b200 as api  # product host mirror (C++ behind the C ABI)
    return api


def _oracle_api():
    from oracle import oracle as api  # CPU restatement (test infrastructure)
    return api


@pytest.fixture(params=["oracle", pytest.param("host", marks=pytest.mark.gpu)])
def api(request):
    """The reference's public API surface (Compiler, Simulation, gates) for tests that RUN programs: once as
    the CPU oracle, once as the product's host mirror -- whose Simulation.run executes on the GPU (there is
    no CPU execution path in the product), hence the gpu mark on that leg."""
    return _oracle_api() if request.param == "oracle" else _host_api()


@pytest.fixture(params=["oracle", "host"])
def capi(request):
    """Same surface for tests that only lower/compile (host logic; both legs run on the CPU)."""
    return _oracle_api() if request.param == "oracle" else _host_api()
