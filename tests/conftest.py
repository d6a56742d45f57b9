import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import cpu_oracle
    cpu_oracle.build()
    return cpu_oracle


def rel_err(got, ref, floor_rel=1e-30):
    """SURVEY 8d parity gate: max |got-ref| / max(|ref|, floor), floor = floor_rel * max|ref|."""
    import numpy as np
    got = np.asarray(got, dtype=float)
    ref = np.asarray(ref, dtype=float)
    floor = floor_rel * np.max(np.abs(ref)) if ref.size else 0.0
    den = np.maximum(np.abs(ref), floor)
    den = np.where(den == 0.0, 1.0, den)
    return float(np.max(np.abs(got - ref) / den)) if ref.size else 0.0
