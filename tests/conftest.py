import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def engine():
    """The process-wide engine on cuda:0 (GPU tests only)."""
    import automative_rag_b200 as rag

    return rag.get_engine(0)


@pytest.fixture(autouse=True)
def _kernel_choice_back_to_auto(request):
    """GPU tests force kernel families through the shared engine; every test starts and ends on AUTO so the files
    behave the same in one pytest process (how the driver runs them) as one process per file."""
    yield
    if request.node.get_closest_marker("gpu") is None:
        return
    import torch

    if not torch.cuda.is_available():
        return
    import automative_rag_b200 as rag
    from automative_rag_b200 import _ffi

    eng = rag.get_engine(0)
    eng.set_dense_impl(_ffi.RS_DENSE_AUTO)
    eng.set_maxsim_impl(_ffi.RS_MAXSIM_AUTO)
    eng.set_scan_trace(None)
