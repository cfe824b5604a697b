"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI symbol checks, gloo.
`-m gpu` runs on a B200: the parity tests proper, through the C-ABI shared library.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "reference: needs /root/reference (authoring container only)")


def pytest_collection_modifyitems(config, items):
    has_ref = Path("/root/reference/src").exists()
    skip_ref = pytest.mark.skip(reason="/root/reference not present on this box")
    for item in items:
        if "reference" in item.keywords and not has_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def golden_mode_sep():
    return np.load(GOLDEN / "mode_sep_fixture.npz", allow_pickle=False)


@pytest.fixture(scope="session")
def golden_latent():
    return np.load(GOLDEN / "latent_ode_fixture.npz", allow_pickle=False)


@pytest.fixture(scope="session")
def golden_rhs():
    return np.load(GOLDEN / "rhs_fixture.npz", allow_pickle=False)
