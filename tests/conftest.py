import gzip
import os
import shutil
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(HERE, "golden")
GOLDEN_DATA = os.path.join(GOLDEN, "data")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def data_dir():
    return GOLDEN_DATA


@pytest.fixture(scope="session")
def real_trees(tmp_path_factory):
    """The two real KO edge lists of the reference, unpacked from the gzipped fixtures."""
    out = tmp_path_factory.mktemp("real_trees")
    paths = {}
    for f in os.listdir(GOLDEN_DATA):
        if f.endswith(".gz"):
            dst = os.path.join(out, f[:-3])
            with gzip.open(os.path.join(GOLDEN_DATA, f), "rb") as src, open(dst, "wb") as d:
                shutil.copyfileobj(src, d)
            paths[f[:-3]] = dst
    return paths
