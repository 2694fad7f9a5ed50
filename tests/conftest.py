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
    from oracle import oracle as orc

    orc.build()
    return orc


def _have_sm100() -> bool:
    try:
        import strling_b200 as sb

        sb.StrGpu(0).close()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `pytest tests` on a box without a B200: gpu-marked tests are skipped instead of failing with STRGPU_ERR_NO_DEVICE
    # (the library has no CPU fallback).  On the GPU box nothing is skipped.
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items or _have_sm100():
        return
    skip = pytest.mark.skip(reason="no sm_100 CUDA device: the scan / cluster kernels have no CPU fallback")
    for it in gpu_items:
        it.add_marker(skip)
