"""CPU, only where the reference tree is mounted (skipped on the GPU box): the German-credit loader reproduces the design
matrix the reference ships (data/germancredit.RData, built by inst/reproducegermancredit/germancredit.preparedata.R)."""
import os

import numpy as np
import pytest

CSV = "/root/reference/inst/reproducegermancredit/german_credit.csv"
RDATA = "/root/reference/data/germancredit.RData"


@pytest.mark.skipif(not (os.path.exists(CSV) and os.path.exists(RDATA)), reason="reference tree not mounted")
def test_german_credit_loader_matches_the_packaged_design():
    from tests.rdx_reader import read_rdata
    from unbiasedmcmc_b200.datasets import german_credit
    Y, X, names = german_credit(CSV)
    ref = read_rdata(RDATA)
    Xr, attrs = ref["X"]
    Xr = Xr.reshape(tuple(attrs["dim"]), order="F")
    assert X.shape == (1000, 49) == Xr.shape
    assert np.array_equal(Y, np.asarray(ref["Y"], dtype=np.float64))
    assert np.array_equal(X, Xr)
    # R's make.names turns every non-alphanumeric character of the CSV header into a dot
    import re
    assert [re.sub(r"[^A-Za-z0-9]", ".", n) if n != "(Intercept)" else n for n in names] == attrs["dimnames"][1]
