import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.hookimpl(tryfirst=True)
def pytest_cmdline_main(config):
    """The CPU suite spends most of its time in the host emulator (one host thread per emulated CUDA thread: latency-,
    not throughput-bound), so on a machine without a GPU the tests are spread over worker processes when pytest-xdist is
    installed and no -n was given (EDK_TEST_WORKERS=0 keeps one process).  Never on a GPU box: the parity tests share
    one device and its memory."""
    opt = config.option
    if getattr(opt, "numprocesses", "no xdist") is not None or os.environ.get("PYTEST_XDIST_WORKER"):
        return None
    if getattr(opt, "collectonly", False) or getattr(opt, "usepdb", False):
        return None
    want = os.environ.get("EDK_TEST_WORKERS", "")
    try:
        import torch

        on_gpu_box = torch.cuda.is_available()
    except Exception:
        on_gpu_box = False
    n = int(want) if want.isdigit() else (0 if on_gpu_box else min(8, os.cpu_count() or 1))
    if n > 1 and not on_gpu_box:
        opt.numprocesses = n
    return None


def rel_err(a, b):
    """Block-wise Frobenius relative error, the measure SURVEY 8c(iii) fixes."""
    import numpy as np

    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def load_golden(name):
    import numpy as np

    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_timeslices(g):
    """(index into g["E"], timeslice) pairs: fixtures at config 1's own shape keep all inputs but only some results."""
    if "timeslices" in g.files:
        return [(i, int(t)) for i, t in enumerate(g["timeslices"])]
    return [(t, t) for t in range(int(g["latt_size"][3]))]


# the reference's stored fixtures are git-LFS pointers in this checkout; if they are ever
# materialised (sha256 below) the tests additionally run the reference's exact inputs against its stored results.
WEAK_FIELD_SHA256 = {
    "weak_field.lime": "de4de48d80869ed537432e237e7cdd99aae36b31f05e2e6f6b4d0b0bc3b28576",
    "weak_field.eigenvector.input.npy": "a6a84f704a86c563066149704851e7579c6ed809cac6b69eeff7f912f64e39c5",
    "weak_field.elemental.npy": "5a244d9c04a94a809adc45f7eec4019dc36fb98218a8259ce6c64c0ac4225193",
    "weak_field.displacement_elemental.npy": "6b706962020f21f4f6f7d4bda3bfbd6163bdb22c0cadb0751c283af9ee375bf4",
}


def reference_weak_field_files():
    """Paths of the reference's tests/weak_field.* when they are the real files, else None (with the reason)."""
    import hashlib

    root = os.path.join(os.environ.get("EDK_REFERENCE_ROOT", "/root/reference"), "tests")
    paths = {}
    for name, sha in WEAK_FIELD_SHA256.items():
        path = os.path.join(root, name)
        if not os.path.exists(path):
            return None, f"{path} does not exist"
        if os.path.getsize(path) < 1024:
            return None, f"{path} is a git-LFS pointer ({os.path.getsize(path)} bytes)"
        with open(path, "rb") as f:
            if hashlib.sha256(f.read()).hexdigest() != sha:
                return None, f"{path} does not have the sha256 of the reference's LFS pointer"
        paths[name] = path
    return paths, ""
