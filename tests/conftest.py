import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(REPO, "tests", "golden", "tiny_ref.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def tiny_root(tmp_path_factory):
    """The seeded tiny dataset the golden vectors were generated on (tests/golden/make_golden.py)."""
    from llmrec_b200.synth import make_dataset
    root = str(tmp_path_factory.mktemp("tiny")) + "/"
    make_dataset(root, dataset="netflix", n_users=300, n_items=400, n_inter=1500, dims=(32, 64, 96), seed=0)
    return root
