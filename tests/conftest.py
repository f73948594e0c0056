import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.bindings import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference (oracle/_ref); skipped where it was never built."""
    from oracle.bindings import Reference, build_ref
    try:
        build_ref()
    except Exception:
        pass
    if not Reference.available():
        pytest.skip("oracle/_ref/libcoup_ref.so not built (needs /root/reference at build time)")
    return Reference()


@pytest.fixture(scope="session")
def kat_scenarios():
    return json.load(open(os.path.join(GOLDEN, "kat_scenarios.json")))["scenarios"]


@pytest.fixture(scope="session")
def playthrough():
    return json.load(open(os.path.join(GOLDEN, "playthrough_coup.json"), encoding="utf-8"))


@pytest.fixture(scope="session")
def ref_trajectories():
    d = np.load(os.path.join(GOLDEN, "ref_trajectories.npz"))
    return d["actions"], d["offsets"], d["records"]


@pytest.fixture(scope="session")
def ref_tensors():
    d = np.load(os.path.join(GOLDEN, "ref_tensors.npz"))
    return d["picks"], d["info"].astype(np.float32), d["obs"].astype(np.float32)
