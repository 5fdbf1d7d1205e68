"""GPU: the batch driver end to end -- runner.run() (reference runner.py:72-161) reading the reference's CSV input
layout ({name}_hifi / _prior / _mf_hyp / _sf_hyp.csv, null_prior.csv) and writing {prefix}_{algo}_{loss,agent,sample}.csv
with the reference's column order (headers of the reference's own logged files, tests/golden/csv_headers.json)."""
import json
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


def _write_inputs(golden_dir, tmp):
    inp = np.load(os.path.join(golden_dir, "inputs_australia6.npz"))
    name = os.path.join(tmp, "australia6")
    pd.DataFrame(inp["truth"], columns=["X", "Y", "f_H"]).to_csv(f"{name}_hifi.csv", index=False)
    pd.DataFrame(inp["prior"], columns=["X", "Y", "f_prior"]).to_csv(f"{name}_prior.csv", index=False)
    pd.DataFrame([inp["mf_hyp"]], columns=["mu_lo", "s^2_lo", "L_lo", "mu_hi", "s^2_hi", "L_hi", "rho", "noise_lo",
                                           "noise_hi"]).to_csv(f"{name}_mf_hyp.csv", index=False)
    pd.DataFrame([inp["sf_hyp"]], columns=["mu_sf", "s^2_sf", "L_sf", "noise_sf"]).to_csv(f"{name}_sf_hyp.csv", index=False)
    null = os.path.join(tmp, "null_prior.csv")
    with open(null, "w") as f:
        f.write("X,Y,f_prior\n")                      # header only, like Data/null_prior.csv
    return name, null


@pytest.mark.parametrize("n_processors", [1, 2])
def test_runner_csv_round_trip(golden_dir, tmp_path, n_processors):
    from mfgp_coverage_b200 import runner
    with open(os.path.join(golden_dir, "csv_headers.json")) as f:
        ref = json.load(f)
    name, null = _write_inputs(golden_dir, str(tmp_path))
    prefix = os.path.join(str(tmp_path), "out")
    A, T, S = 8, 4, 3
    runner.run(n_processors=n_processors, name=name, prefix=prefix, agents=A, iterations=T, simulations=S, sigma_n=0.1,
               algorithms=["todescato_hmf", "todescato_nsf", "lloyd"], seed=5, null_prior_path=null)
    for algo, stem in (("todescato_hmf", "australia6_todescato_hmf"), ("lloyd", "australia6_lloyd"),
                       ("todescato_nsf", "australia6_todescato_hmf")):
        for kind in ("loss", "agent", "sample"):
            path = f"{prefix}_{algo}_{kind}.csv"
            with open(path) as f:
                header = f.readline().rstrip("\n")
            assert header == ref[f"{stem}_{kind}"]["header"], (algo, kind)           # pandas default index column first
        loss = pd.read_csv(f"{prefix}_{algo}_loss.csv", index_col=0)
        agent = pd.read_csv(f"{prefix}_{algo}_agent.csv", index_col=0)
        assert len(loss) == S * T and len(agent) == S * T * A
        assert list(loss.SimNum) == sorted(loss.SimNum) and set(loss.SimNum) == set(range(S))   # concatenated in sim order
        assert list(loss.index) == list(range(S * T))
        assert np.all(np.isfinite(loss.Loss)) and np.all(loss.Loss > 0)
        fid = {"todescato_hmf": "M", "todescato_nsf": "S", "lloyd": None}[algo]
        if fid:
            assert set(loss.Fidelity) == {fid}
            assert np.allclose(agent.Var0, agent.Var0.iloc[0]) and agent.Var0.iloc[0] > 0
        else:
            assert loss.Fidelity.isna().all()              # "NA" strings read back as NaN, as in the reference's files
    smp = pd.read_csv(f"{prefix}_todescato_hmf_sample.csv", index_col=0)
    assert len(smp) > 0 and smp.Iteration.min() >= 1       # iteration 0 never samples (simulator.py:858)
    # the same seeds give the same CSVs whatever the number of worker processes (run-sharding is order-preserving)
    if n_processors == 2:
        runner.run(n_processors=1, name=name, prefix=prefix + "_single", agents=A, iterations=T, simulations=S, sigma_n=0.1,
                   algorithms=["lloyd"], seed=5, null_prior_path=null)
        a = pd.read_csv(f"{prefix}_lloyd_loss.csv", index_col=0)
        b = pd.read_csv(f"{prefix}_single_lloyd_loss.csv", index_col=0)
        assert a.shape == b.shape
