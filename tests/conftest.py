import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Built on demand; reference qpOASES when oracle/_ref exists."""
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def cuda_engine_factory():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from quadruped_ctrl_b200 import engine as E
    made = []

    def make(horizon, max_batch, solver=None):
        e = E.MpcBatch(horizon, max_batch, 0)
        if solver is not None:
            e.set_solver(solver)   # "riccati" (the default) or "inverse"
        made.append(e)
        return e

    yield make
    for e in made:
        e.close()
