import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def scenario_dir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("scenarios"))


def _make(scenario_dir, rows, cols, name, dense=None, **cfg):
    from cityflow_b200 import scenario
    return scenario.make_grid_scenario(scenario_dir, rows, cols, dense=dense, name=name, **cfg)


@pytest.fixture(scope="session")
def cfg_1x1(scenario_dir):
    return _make(scenario_dir, 1, 1, "g1")


@pytest.fixture(scope="session")
def cfg_3x3_dense(scenario_dir):
    return _make(scenario_dir, 3, 3, "g3d", dense=dict(frac=1.0, interval=3.0, seed=3))


@pytest.fixture(scope="session")
def cfg_6x6(scenario_dir):
    return _make(scenario_dir, 6, 6, "g6")


@pytest.fixture(scope="session")
def cfg_6x6_dense(scenario_dir):
    return _make(scenario_dir, 6, 6, "g6d", dense=dict(frac=1.0, interval=4.0, seed=1))


@pytest.fixture(scope="session")
def cfg_6x6_rl(scenario_dir):
    return _make(scenario_dir, 6, 6, "g6rl", dense=dict(frac=1.0, interval=5.0, seed=2), rl_traffic_light=True)
