"""north_star (a): 'a replay mode that feeds the reference's own recorded uniform/normal draws into both
implementations'.  The draws are those of the R session behind the reference's README (set.seed(1), emulated stream of
oracle/r_stream.hpp, recorded per replicate as typed tapes); the CUDA engine replays them through the C ABI and must
return the chains of that session bit for bit — and therefore the literals R printed (README.md:146, :171)."""
import numpy as np
import pytest

from tests.test_r_stream import readme_session
from unbiasedmcmc_b200 import _abi as A
from unbiasedmcmc_b200 import registry as R

pytestmark = pytest.mark.gpu


def test_readme_session_replayed_on_the_device(engine, oracle):
    mdl, mt, ch, s = readme_session(oracle, record=True)
    tu, tn, _ = s.tapes
    draws = R.Draws(tape_u=np.nan_to_num(tu), tape_n=np.nan_to_num(tn))
    g = engine.sample_coupled_chains(mdl, draws, m=2000, lag=500, stride=2101, nrep=500)
    for key in ("meetingtime", "iteration", "cost"):
        assert np.array_equal(g[key], ch[key]), key
    for r in range(500):
        n = int(ch["iteration"][r]) + 1
        assert np.array_equal(g["samples1"][r, :n], ch["samples1"][r, :n])
        assert np.array_equal(g["samples2"][r, :n - 500], ch["samples2"][r, :n - 500])
    assert f"{g['cost'].mean():.7g}" == "2072.788"                                  # README.md:146
    est = engine.h_bar(g, A.UBM_H_IDENTITY, k=500, m=2000)[:, 0]
    assert f"{est.mean():.7g}" == "0.005471002"                                     # README.md:171
    assert f"{1.96 * est.std(ddof=1) / np.sqrt(500):.7g}" == "0.1458675"
    # the on-the-fly estimator of the same replayed session (R/unbiasedestimator.R:43-104) agrees with H_bar
    ue = engine.sample_unbiasedestimator(mdl, draws, k=500, m=2000, lag=500, nrep=500)
    assert np.array_equal(ue["meetingtime"], ch["meetingtime"])
    np.testing.assert_allclose(ue["uestimator"][:, 0], est, rtol=1e-12, atol=1e-12)
