import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _strict_fp32():
    """The unfused drop-in path runs torch's own 1x1 conv (cuDNN); TF32 is on by default for
    convolutions and is worth ~2e-3 relative error, more than the fp32 tolerance (1e-3) the
    parity tests apply.  The fused kernels never use TF32."""
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def ref_modules_golden(golden_dir):
    import numpy as np
    return dict(np.load(os.path.join(golden_dir, "ref_modules.npz")))


@pytest.fixture(scope="session")
def ref_cuda_golden(golden_dir):
    import numpy as np
    path = os.path.join(golden_dir, "ref_cuda_ops.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/ref_cuda_ops.npz not generated yet (needs the GPU box)")
    return dict(np.load(path))
