import importlib
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN_DIR = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def pytest_collection_modifyitems(config, items):
    # GPU tests never run silently on a box without a GPU: they are skipped with a reason unless
    # selected on a machine that has one.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("tce-rvos_b200")


@pytest.fixture(scope="session")
def capi():
    return importlib.import_module("tce-rvos_b200.capi")


@pytest.fixture(scope="session")
def oracle():
    from oracle import msda_oracle
    msda_oracle.build_c_oracle()
    return msda_oracle


def golden_names():
    return sorted(p.stem for p in GOLDEN_DIR.glob("*.npz"))


def load_golden(name):
    import numpy as np
    return dict(np.load(GOLDEN_DIR / f"{name}.npz"))
