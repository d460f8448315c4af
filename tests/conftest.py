import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["c1_n13_dmso", "mixed_n21", "c2_n50_water", "mixed_n50", "n100_mixed", "sulfur_n30", "epsband_n21"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run under gpurun on a B200)")
    # a fresh clone has no built libraries (they are git-ignored): build them once; failures surface in the tests
    try:
        from gnnimplicitsolvent_b200 import _lib, build, torch_op
        if not os.path.exists(_lib.LIB_PATH) or not os.path.exists(torch_op.LIB_PATH):
            build.build_library()
            build.build_torch_op()
    except Exception as exc:   # noqa: BLE001
        print(f"conftest: could not build the native libraries: {exc}", file=sys.stderr)


def load_case(name):
    return dict(np.load(os.path.join(GOLDEN, f"case_{name}.npz")))


def load_weights():
    import torch
    z = np.load(os.path.join(GOLDEN, "gnn_weights.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="session")
def weights():
    return load_weights()


def rel_rms(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(((a - b) ** 2).sum()) / np.sqrt((b ** 2).sum()))
