import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) on a box without a CUDA device or without the built library."""
    have = False
    try:
        import torch

        from car_trailer_mpc_b200 import _lib

        have = torch.cuda.is_available() and os.path.exists(_lib.lib_path())
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and the built libttmpc.so")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def traj():
    from car_trailer_mpc_b200 import problem as pb

    return pb.load_reference_trajectory()


@pytest.fixture(scope="session")
def cfg40():
    from car_trailer_mpc_b200 import tracking_preset

    return tracking_preset(40)
