import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracles():
    from oracle import cpu_oracle
    cpu_oracle.build()
    return {n: cpu_oracle.CurveOracle(n) for n in ("bn254", "bls12_381", "bn254_g2", "bls12_381_g2")}
