import os
import sys

import pytest

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


def pytest_sessionfinish(session, exitstatus):
    """Observed c_high / RTVQ-code equality rates of the parity runs of this session -> gpurun_out/parity_rates.json
    (copied to profiles/parity_rates.json when committed)."""
    try:
        from tests import parity
    except Exception:
        return
    if not parity.RATES:
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    tot = {k: sum(r[k] for r in parity.RATES) for k in ("c_high_equal", "c_high_total", "codes_equal", "codes_total")}
    with open(os.path.join(out, "parity_rates.json"), "w") as f:
        json.dump({"totals": tot, "max_absolute_error_rel_dev_fp16_bases": parity.MAX_ABS_DEV, "runs": parity.RATES}, f, indent=1)
