import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def fixtures():
    return load_golden("fixtures_testsupply.json")


@pytest.fixture(scope="session")
def small_mixed(fixtures):
    from serenity_b200.inputs.basis import shell_table_from_list
    return shell_table_from_list(fixtures["bases"]["SMALL_MIXED"])


def grid_arrays(fixtures, name):
    import numpy as np
    g = fixtures["grids"][name]
    xyz = np.stack([g["x"], g["y"], g["z"]], axis=1).astype(np.float64)
    return np.ascontiguousarray(xyz), np.asarray(g["w"], dtype=np.float64)
