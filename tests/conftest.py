import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")



@pytest.fixture(scope="session")
def cov():
    import coverage_b200
    return coverage_b200


@pytest.fixture(scope="session")
def orc():
    from oracle import c_oracle
    c_oracle.build()
    return c_oracle


@pytest.fixture(scope="session")
def npo():
    from oracle import coverage_oracle
    return coverage_oracle


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def fire_rows(cov):
    return cov.fire_io.load_fire_rows_npz(os.path.join(GOLDEN, "fire_rows.npz"))


@pytest.fixture(scope="session")
def targets():
    with np.load(os.path.join(GOLDEN, "targets.npz")) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture()
def engine(cov):
    e = cov.CoverageEngine(0)
    yield e
    e.close()
