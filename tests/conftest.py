import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def rel_err(a, b):
    """Block-wise Frobenius relative error, the measure SURVEY 8c(iii) fixes."""
    import numpy as np

    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def load_golden(name):
    import numpy as np

    return np.load(os.path.join(GOLDEN, name + ".npz"))
