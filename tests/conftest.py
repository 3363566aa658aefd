import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (reference Cython in oracle/_ref + NumPy glue).  Builds oracle/_ref from
    /root/reference when that tree is present; on the GPU box the prebuilt files are used."""
    from oracle import build_ref, ref_glue

    if not build_ref.build(verbose=False):
        pytest.fail("oracle/_ref is not built and /root/reference is absent")
    ref_glue.ref()
    return ref_glue


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)
