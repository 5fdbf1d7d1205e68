"""Hyper-parameter training (reference gaussian_process.py:81-119, :344-399, trainer.py:17-92).
CPU: the oracle's NLML + analytic gradient against the LIVE reference's `likelihood` values and finite differences of them
(tests/golden/ref_train.npz, written by oracle/make_golden.py).  GPU: mfgp_nlml_grad and the train() loop against both."""
import os

import numpy as np
import pytest

from oracle import train as otrain

E2 = np.empty((0, 2))
E1 = np.empty((0, 1))


def _g(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_train.npz"))


def test_oracle_nlml_and_gradient_vs_live_reference(golden_dir):
    g = _g(golden_dir)
    for k in range(3):
        v, gr = otrain.nlml_and_grad(g["mf_hyps"][k], g["X_L"], g["y_L"], g["X_H"], g["y_H"])
        assert abs(v - g["mf_nlml"][k]) <= 1e-12 * max(1.0, abs(v))
        assert np.max(np.abs(gr - g["mf_grad_fd"][k])) <= 1e-7 * np.max(np.abs(gr))
        v, gr = otrain.nlml_and_grad(g["sf_hyps"][k], E2, E1, g["X_H"], g["y_H"])
        assert abs(v - g["sf_nlml"][k]) <= 1e-12 * max(1.0, abs(v))
        assert np.max(np.abs(gr - g["sf_grad_fd"][k])) <= 1e-7 * np.max(np.abs(gr))
    with pytest.raises(TypeError):
        otrain.nlml_and_grad(np.zeros(5), E2, E1, g["X_H"], g["y_H"])


def test_oracle_training_lowers_the_nlml(golden_dir):
    g = _g(golden_dir)
    h0 = g["mf_hyps"][0]
    res = otrain.train(h0, g["X_L"][:60], g["y_L"][:60], g["X_H"][:60], g["y_H"][:60], maxiter=25)
    assert res.fun < otrain.nlml(h0, g["X_L"][:60], g["y_L"][:60], g["X_H"][:60], g["y_H"][:60]) - 10.0


@pytest.mark.gpu
def test_device_nlml_and_gradient(golden_dir):
    from mfgp_coverage_b200.gaussian_process import MFGP, SFGP
    g = _g(golden_dir)
    m = MFGP(g["X_L"], g["y_L"], g["X_H"], g["y_H"], 0.5, 0.1)
    assert np.array_equal(m.hyp, g["mf_hyps"][0])                     # the constructor's initial hyper-parameters (:297-312)
    s = SFGP(g["X_H"], g["y_H"], 0.01)
    assert np.array_equal(s.hyp, g["sf_hyps"][0])
    for k in range(3):
        for model, hyps, key, data in ((m, g["mf_hyps"], "mf", (g["X_L"], g["y_L"], g["X_H"], g["y_H"])),
                                       (s, g["sf_hyps"], "sf", (E2, E1, g["X_H"], g["y_H"]))):
            v, gr = model.likelihood_and_grad(hyps[k])
            vo, go = otrain.nlml_and_grad(hyps[k], *data)
            assert abs(v - g[f"{key}_nlml"][k]) <= 1e-9 * max(1.0, abs(vo)), (key, k)       # the live reference's value
            assert abs(v - vo) <= 1e-9 * max(1.0, abs(vo))
            assert gr.shape == go.shape and np.max(np.abs(gr - go)) <= 1e-7 * np.max(np.abs(go)), (key, k)
            assert model.likelihood(hyps[k]) == v
    assert np.array_equal(m.hyp, g["mf_hyps"][0])                     # evaluating the likelihood leaves `hyp` alone
    # odd sizes (padding rows of the 64-tiles must not leak into the sums) and a model that is also used for prediction
    m2 = MFGP(g["X_L"][:37], g["y_L"][:37], g["X_H"][:91], g["y_H"][:91], 0.5, 0.1)
    v, gr = m2.likelihood_and_grad(g["mf_hyps"][1])
    vo, go = otrain.nlml_and_grad(g["mf_hyps"][1], g["X_L"][:37], g["y_L"][:37], g["X_H"][:91], g["y_H"][:91])
    assert abs(v - vo) <= 1e-9 * max(1.0, abs(vo)) and np.max(np.abs(gr - go)) <= 1e-7 * np.max(np.abs(go))
    with pytest.raises(TypeError):
        m.likelihood(np.zeros(4))


@pytest.mark.gpu
def test_device_training_follows_the_oracle_optimiser(golden_dir, tmp_path, capsys):
    from mfgp_coverage_b200 import trainer
    from mfgp_coverage_b200.gaussian_process import MFGP
    g = _g(golden_dir)
    X_L, y_L, X_H, y_H = g["X_L"][:100], g["y_L"][:100], g["X_H"][:100], g["y_H"][:100]
    m = MFGP(X_L, y_L, X_H, y_H, 0.5, 0.1)
    h0 = m.hyp.copy()
    res = m.train(maxiter=8)
    ref = otrain.train(h0, X_L, y_L, X_H, y_H, maxiter=8)
    assert "Log likelihood" in capsys.readouterr().out                # reference callback (:219-227) prints every step
    assert res.fun < otrain.nlml(h0, X_L, y_L, X_H, y_H) - 10.0
    assert abs(res.fun - ref.fun) <= 1e-5 * max(1.0, abs(ref.fun))    # same optimiser, gradients equal to ~1e-9
    assert np.max(np.abs(res.x - ref.x)) <= 1e-3
    assert np.array_equal(m.hyp, res.x)
    # trainer.py round trip with the reference's file layout and column labels (trainer.py:27-52, :66-92)
    import pandas as pd
    d = str(tmp_path)
    pd.DataFrame(np.column_stack((X_L, y_L)), columns=["X", "Y", "f_L"]).to_csv(f"{d}/t_lofi_train.csv", index=False)
    pd.DataFrame(np.column_stack((X_H, y_H)), columns=["X", "Y", "f_H"]).to_csv(f"{d}/t_hifi_train.csv", index=False)
    trainer.train_mfgp("t", data_dir=d, save=True, callback=False)
    trainer.train_sfgp("t", data_dir=d, save=True, callback=False)
    mf = pd.read_csv(f"{d}/t_mf_hyp.csv")
    sf = pd.read_csv(f"{d}/t_sf_hyp.csv")
    assert list(mf.columns) == trainer.MF_LABELS and list(sf.columns) == trainer.SF_LABELS and len(mf) == len(sf) == 1
    v_trained = otrain.nlml(mf.values[0], X_L, y_L, X_H, y_H)
    assert v_trained < otrain.nlml(h0, X_L, y_L, X_H, y_H) - 10.0
    # the written files are what simulator.init_MFGP / init_SFGP consume
    from mfgp_coverage_b200 import simulator as sim
    assert sim._fidelity_of(mf) == "M" and sim._fidelity_of(sf) == "S"
