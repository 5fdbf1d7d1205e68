"""CPU, build container only: the oracle against the UNMODIFIED reference imported live from /root/reference
(skipped on the GPU box, where the reference does not exist and the committed fixtures take over)."""
import random

import numpy as np
import pytest

from oracle import algorithms as oalg
from oracle import gp as ogp
from oracle import reference_live as rl
from tests import synth

pytestmark = pytest.mark.skipif(not rl.available(), reason="/root/reference not present")


def test_predict_small_grid_live():
    sim, gp = rl.load()
    import pandas as pd
    xy = synth.grid(21)
    f = synth.truth_function(xy)
    X_L, y_L, X_H, y_H = synth.training_set(xy, f, 60)
    hyp = pd.DataFrame([synth.MF_HYP])
    m = sim.init_MFGP(hyp, None)
    m.updt_info(X_L, y_L, X_H, y_H)
    mu_r, cov_r = m.predict(xy)
    om = ogp.Model(ogp.GPParams.from_hyp(synth.MF_HYP), X_L, y_L, X_H, y_H)
    om.updt_info()
    mu, var = om.predict(xy)
    assert np.max(np.abs(mu - mu_r[:, 0])) <= 1e-12 and np.max(np.abs(var - np.diag(cov_r))) <= 1e-14
    mu_e, var_e = om.predict(xy, exact_solve=True)          # same LAPACK calls as the reference
    assert np.max(np.abs(mu_e - mu_r[:, 0])) <= 1e-14 and np.max(np.abs(var_e - np.diag(cov_r))) <= 1e-15


def test_todescato_loop_live():
    sim, gp = rl.load()
    import pandas as pd
    xy = synth.grid(21)
    truth_arr = np.column_stack((xy, synth.truth_function(xy)))
    truth = pd.DataFrame(truth_arr, columns=["X", "Y", "f_H"])
    hyp = pd.DataFrame([synth.SF_HYP])
    prior = pd.DataFrame(np.empty((0, 3)), columns=["X", "Y", "f_prior"])
    rl.seeded(sim, 5)
    pos = np.column_stack(([random.random() for _ in range(4)], [random.random() for _ in range(4)]))
    l1, a1, s1 = sim.todescato("t", 0, 8, 4, pos.copy(), truth, 0.1, prior, hyp, False, None, True)
    r = random.Random(5)
    for _ in range(8):
        r.random()
    l2, a2, s2 = oalg.todescato(0, 8, 4, pos.copy(), truth_arr, 0.1, None, synth.SF_HYP, r, np.random.default_rng(5))
    assert max(abs(x["Loss"] - y["Loss"]) for x, y in zip(l1, l2)) <= 1e-14
    assert max(abs(x["XCentroid"] - y["XCentroid"]) for x, y in zip(a1, a2)) <= 1e-13
    assert len(s1) == len(s2)
