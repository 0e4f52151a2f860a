"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` covers the oracle against the golden vectors / the compiled reference, the host logic
and the C-ABI surface (no compute calls); `-m gpu` holds the parity tests proper, which call through
the C ABI on a real B200.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle_lib import Reference, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref/libgpsref.so not available (no /root/reference and no prebuilt copy)")
    return Reference()


@pytest.fixture(scope="session")
def golden():
    path = REPO / "tests" / "golden" / "golden_l2.npz"
    if not path.exists():
        pytest.skip("golden fixtures missing; run tests/golden/make_golden.py in the authoring container")
    return np.load(path)


@pytest.fixture(scope="session")
def engine():
    """One engine context on cuda:0 for the whole GPU test session."""
    from stm32f4_sdr_gps_b200 import Engine
    eng = Engine(device=0, max_sv=40, ring_ms=256)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def host_engine():
    """Engine sized for the host-side mirror: satellite slot == PRN (1..210), 1 s signal ring."""
    from stm32f4_sdr_gps_b200 import Engine, load_host_library
    eng = Engine(device=0, max_sv=211, ring_ms=1024)
    lib = load_host_library()
    assert lib.gpsb_host_attach(eng.handle) == 0
    yield eng
    lib.gpsb_host_attach(None)
    eng.close()
