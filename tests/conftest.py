"""pytest configuration: the `gpu` marker and import path.

`python -m pytest tests -m "not gpu"` runs anywhere (oracle vs golden vectors, host logic, C-ABI symbol check);
`python -m pytest tests -m gpu` needs a B200 and goes through the C ABI of libhpem.so.
"""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200, sm_100a) and the built libhpem.so')


@pytest.fixture(scope='session')
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.fail('a test marked `gpu` was collected on a machine without CUDA')
    return 0
