"""CPU: the oracle against its own committed known-answer vectors (tests/golden/oracle_kat.json).
Bit-exact pieces are pinned by SHA-256, floating-point pieces by stored values."""
import importlib.util
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _compute():
    spec = importlib.util.spec_from_file_location("make_kat", os.path.join(HERE, "golden", "make_kat.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.compute()


def test_oracle_known_answers():
    want = json.load(open(os.path.join(HERE, "golden", "oracle_kat.json")))
    got = _compute()
    assert set(got) == set(want)
    for k, v in want.items():
        if isinstance(v, str):
            assert got[k] == v, k  # integer / byte outputs: bit-exact
        else:
            np.testing.assert_allclose(np.array(got[k]), np.array(v), rtol=1e-6, atol=1e-7, err_msg=k)
