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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def kat():
    """the reference's own golden vectors (oracle/extract_golden.py)"""
    with open(os.path.join(GOLDEN, "ref_kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def refout():
    """outputs of the compiled reference on seeded inputs (oracle/gen_golden.py)"""
    return dict(np.load(os.path.join(GOLDEN, "ref_outputs.npz")))


@pytest.fixture(scope="session")
def refrn():
    """root-Nyquist Kaiser prototypes and channelizers built on them, from the compiled reference
    (oracle/gen_golden_next.py)"""
    return dict(np.load(os.path.join(GOLDEN, "ref_outputs_next.npz")))


@pytest.fixture(scope="session")
def refrn2():
    """resamp2_crcf / msresamp2_crcf outputs of the compiled reference (oracle/gen_golden_next2.py)"""
    return dict(np.load(os.path.join(GOLDEN, "ref_outputs_next2.npz")))


@pytest.fixture(scope="session")
def refr2():
    """fixtures added in round 2 from the compiled reference (oracle/gen_golden_r2.py)"""
    return dict(np.load(os.path.join(GOLDEN, "ref_outputs_r2.npz")))


@pytest.fixture(scope="session")
def orc():
    """the CPU oracle (checker only)"""
    from oracle import oracle as O
    O.build(ref=os.path.isdir("/root/reference"))
    return O


def cplx(pairs):
    a = np.asarray(pairs, dtype=np.float64)
    return (a[:, 0] + 1j * a[:, 1]).astype(np.complex64)


def rel_rms_per_channel(y, yref):
    """north_star parity metric: per channel k, sqrt(sum_t |Y-Yref|^2 / sum_t |Yref|^2)"""
    y = np.asarray(y, np.complex128)
    yref = np.asarray(yref, np.complex128)
    num = (np.abs(y - yref) ** 2).sum(axis=0)
    den = (np.abs(yref) ** 2).sum(axis=0)
    return np.sqrt(num / np.maximum(den, 1e-300))


def max_abs_norm(y, yref, x):
    """north_star parity metric: max |Y-Yref| / sqrt(mean |x|^2)"""
    pin = np.sqrt(np.mean(np.abs(np.asarray(x, np.complex128)) ** 2))
    return np.abs(np.asarray(y, np.complex128) - np.asarray(yref, np.complex128)).max() / pin


# tolerances stated by BASELINE.json north_star
TOL_REL_RMS = 1e-5
TOL_MAX_ABS = 1e-4
