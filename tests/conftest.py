import os
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))
GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_report_header(config):
    """Which box / GPU the suite ran on (read back when a GPU-side result differs between runs)."""
    try:
        import torch
        if not torch.cuda.is_available():
            return "svr_b200: no CUDA device"
        p = torch.cuda.get_device_properties(0)
        return (f"svr_b200: {p.name} sms={p.multi_processor_count} mem={p.total_memory >> 30}GiB uuid={p.uuid} "
                f"host={os.uname().nodename} torch={torch.__version__} cudnn={torch.backends.cudnn.version()}")
    except Exception as e:      # never let the header break the run
        return f"svr_b200: device query failed ({e})"


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    class _G:
        def __getitem__(self, name):
            return np.load(GOLDEN / f"{name}.npz")
    return _G()
