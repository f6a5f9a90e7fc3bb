import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def traj():
    from car_trailer_mpc_b200 import problem as pb

    return pb.load_reference_trajectory()


@pytest.fixture(scope="session")
def cfg40():
    from car_trailer_mpc_b200 import tracking_preset

    return tracking_preset(40)
