import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")
FIXTURES = os.path.join(GOLDEN, "ref_fixtures")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def fuzz_cases():
    """Golden single-window cases produced by the reference's Sstar.compute (make_golden.py)."""
    z = np.load(os.path.join(GOLDEN, "sstar_fuzz.npz"))
    meta = z["meta"]
    cases, ro, to, po, uo = [], 0, 0, 0, 0
    for V, R, T, b, mx, p in meta:
        ref = z["ref"][ro:ro + V * R].reshape(V, R); ro += V * R
        tgt = z["tgt"][to:to + V * T].reshape(V, T); to += V * T
        pos = z["pos"][po:po + V]; po += V
        cases.append(dict(ref=ref, tgt=tgt, pos=pos, params=(int(b), int(mx), int(p)),
                          score=z["score"][uo:uo + T], nsnp=z["nsnp"][uo:uo + T]))
        uo += T
    return cases


@pytest.fixture(scope="session")
def test0_golden():
    return np.load(os.path.join(GOLDEN, "test0_windows.npz"))
