import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "speech-decoding_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


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
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _exact_fp32_reference():
    """The PyTorch statements the GPU tests compare against must be true fp32 (cuDNN/cuBLAS default to
    TF32 for convolutions, which is 1e-3-level noise and would mask real errors)."""
    try:
        import torch
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    except Exception:
        pass
    yield


# ---- measured parity margins ------------------------------------------------------------------------------
# Every parity test records the worst error it measured (and the tolerance it was held to) through
# tests.parity_log.record(); at session end the table is written to gpurun_out/parity.json (when that directory
# exists, i.e. on the GPU box) so the measured margins -- not only pass/fail -- are committed under profiles/.
def pytest_sessionfinish(session, exitstatus):
    try:
        from tests import parity_log
        parity_log.dump(os.path.join(ROOT, "gpurun_out"))
    except Exception:
        pass
