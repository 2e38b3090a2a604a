import os
import sys

import pytest

# no checkpoints / CLIP caches exist on the box: the tests opt in to seeded synthetic weights (core.synthetic_or_raise)
os.environ.setdefault("SDB_SYNTHETIC_WEIGHTS", "1")
os.environ.setdefault("SDB_NO_TRIAL_DIRS", "1")  # load_config() would create outputs/<name>/<tag>@<timestamp>/ otherwise

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
