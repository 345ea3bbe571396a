import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds (still part of the default run)")


@pytest.fixture(scope="session")
def small_problem():
    from dpgo_ros_b200 import datasets
    return datasets.load_g2o_problem("smallGrid3D", 2)


@pytest.fixture(scope="session")
def tiny_problem():
    from dpgo_ros_b200 import datasets
    return datasets.load_g2o_problem("tinyGrid3D", 2)


@pytest.fixture(scope="session")
def sphere8_problem():
    from dpgo_ros_b200 import datasets
    return datasets.load_g2o_problem("sphere2500", 8)
