"""Golden regression vectors (tests/golden/oracle_r01.json, made by tests/golden/make_golden.py from the oracle): the
oracle (CPU) and the CUDA engine (GPU) must both reproduce them bit for bit.  They pin the engine's own canonical
arithmetic between rounds (parity with the reference is pinned by tests/test_r_stream.py and tests/test_ref_parity.py)."""
import importlib.util
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)
GOLD = json.load(open(os.path.join(HERE, "golden", "oracle_r01.json")))


def _check(backend, kw_est):
    for name, fn in mg.golden_cases().items():
        got = mg.record(fn(backend, kw_est if "estimator" in name else {}))
        assert got.keys() == GOLD[name].keys(), name
        for key in got:
            assert got[key] == GOLD[name][key], (name, key)


def test_oracle_reproduces_golden_vectors(oracle):
    _check(oracle, {"nthreads": 2})
    # the fixtures are not degenerate
    assert any(x != GOLD["c1_estimator"]["uestimator"][0] for x in GOLD["c1_estimator"]["uestimator"])
    assert all(np.isfinite(float.fromhex(x)) for x in GOLD["c2_estimator_d9"]["uestimator"])


@pytest.mark.gpu
def test_engine_reproduces_golden_vectors(engine):
    _check(engine, {})
