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
    """CPU oracle (test infrastructure): compiled on first use."""
    import torch
    from oracle import ts_oracle
    ts_oracle.build()
    # torch's fp32 CPU kernels give thread-count dependent whole-model gradients (2e-3 .. 9e-3 from the fp64 result at
    # 1, 3, 12 or 16 intra-op threads, ~5e-6 at 4-8: scripts/smoke_repeat.py), so the checker runs with
    # at most 8 threads
    torch.set_num_threads(max(1, min(8, torch.get_num_threads())))
    return ts_oracle


@pytest.fixture(scope="session")
def cuda_lib():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test selected but no CUDA device is visible")
    from u2mkd_b200 import _lib
    return _lib.lib()
