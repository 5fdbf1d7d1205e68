"""GPU: the batched replicate stepper (csrc/batched.cu, BASELINE config 5) -- many independent runs advanced together on the
device -- against the oracle loops run one simulation at a time with the same start positions and random streams, and the
runner's batched mode against its sequential mode."""
import os
import random

import numpy as np
import pandas as pd
import pytest

from oracle import algorithms as oalg
from tests import synth
from tests.test_gpu_algorithms import _agent_array, _compare

pytestmark = pytest.mark.gpu


def _inputs(n, multi):
    xy = synth.grid(n)
    truth_arr = np.column_stack((xy, synth.truth_function(xy)))
    lat_xy = np.random.default_rng(99).random((9, 2))      # asymmetric prior: no bit-exact variance ties
    near = np.argmin(((xy[None, :, :] - lat_xy[:, None, :]) ** 2).sum(axis=2), axis=1)
    prior_arr = np.column_stack((lat_xy, 0.8 * truth_arr[near, 2] + 0.02))
    return truth_arr, (prior_arr if multi else None), (synth.MF_HYP if multi else synth.SF_HYP)


def _rows_to_arrays(logs, A):
    lg, ag, sg = logs
    samples = np.array([[float(r["Iteration"]), float(r["Agent"]), float(r["X"]), float(r["Y"]), float(r["Sample"])]
                        for r in sg if r["Agent"] != "NA"]).reshape(-1, 5)
    return np.array([r["Loss"] for r in lg]), _agent_array(ag, len(lg), A), samples


@pytest.mark.parametrize("algo,n,A,T,multi", [("periodic", 32, 5, 24, True), ("todescato", 40, 6, 20, False),
                                              ("todescato", 36, 4, 18, True), ("lloyd", 32, 7, 15, False)])
def test_batched_runs_match_the_oracle_loops_run_by_run(algo, n, A, T, multi):
    from mfgp_coverage_b200 import simulator as sim
    truth_arr, prior_arr, hyp = _inputs(n, multi)
    R = 4
    starts = np.stack([synth.agents(A, 300 + r) for r in range(R)])
    # the uniforms exactly as a sequential todescato run draws them: one per agent and iteration, iteration-major
    streams = [random.Random(500 + r) for r in range(R)]
    unif = np.array([[[s.random() for _ in range(A)] for _ in range(T)] for s in streams])
    pos = starts.copy()
    logs = sim.run_batched(algo, list(range(R)), T, A, pos, truth_arr, 0.1, prior_arr, hyp,
                           uniforms=unif if algo == "todescato" else None,
                           noise_rngs=[np.random.default_rng(700 + r) for r in range(R)], exact_tie_loss=True)
    assert len(logs) == R
    for r in range(R):
        p0 = starts[r].copy()
        if algo == "lloyd":
            ref = oalg.lloyd(r, T, A, p0, truth_arr)
        else:
            ref = getattr(oalg, algo)(r, T, A, p0, truth_arr, 0.1, prior_arr, hyp, random.Random(500 + r), np.random.default_rng(700 + r))
        l, a, s = _rows_to_arrays(logs[r], A)
        lo, ao, so = _rows_to_arrays(ref, A)
        _compare(l, a, s, lo, ao, so, truth_arr)
        assert [row["SimNum"] for row in logs[r][0]] == [r] * T
        if algo != "lloyd":
            assert np.allclose(pos[r], p0, atol=1e-9)          # start positions are advanced in place, like the reference's loops
    if algo == "lloyd":
        assert logs[0][2][0]["Agent"] == "NA" and len(logs[0][2]) == T


def test_runner_batched_mode_writes_the_same_csvs(golden_dir, tmp_path):
    """runner.run(batched=True): the same random streams as the sequential loop (start positions, todescato's uniforms), so
    with a fixed seed the deterministic outputs coincide: everything for lloyd; for todescato everything that does not depend
    on the (unseeded, as in the reference) sample noise, i.e. iteration 0 and the start positions."""
    from mfgp_coverage_b200 import runner
    from tests.test_gpu_runner import _write_inputs
    name, null = _write_inputs(golden_dir, str(tmp_path))
    kw = dict(name=name, agents=8, iterations=6, simulations=5, sigma_n=0.1, seed=9, null_prior_path=null, n_processors=1)
    try:
        runner.run(prefix=os.path.join(str(tmp_path), "seq"), algorithms=["lloyd", "todescato_hmf"], batched=False, **kw)
        runner.run(prefix=os.path.join(str(tmp_path), "bat"), algorithms=["lloyd", "todescato_hmf"], batched=True, **kw)
    finally:
        runner.BATCHED = False
        os.environ["MFGP_BATCHED"] = "0"
    for algo in ("lloyd", "todescato_hmf"):
        for kind in ("loss", "agent"):
            a = pd.read_csv(os.path.join(str(tmp_path), f"seq_{algo}_{kind}.csv"), index_col=0)
            b = pd.read_csv(os.path.join(str(tmp_path), f"bat_{algo}_{kind}.csv"), index_col=0)
            assert list(a.columns) == list(b.columns) and a.shape == b.shape
            if algo != "lloyd":
                a, b = a[a.Iteration == 0], b[b.Iteration == 0]
            num = [c for c in a.columns if c != "Fidelity"]
            assert np.allclose(a[num].values.astype(float), b[num].values.astype(float), rtol=1e-9, atol=1e-12), (algo, kind)


def test_whole_run_replayed_from_one_cuda_graph_is_bitwise_the_eager_run():
    """the device-resident iteration loop (sample -> append -> posterior -> both partitions -> decision ->
    move, simulator.py:888-904) captured once as a CUDA graph: same log buffers, bit for bit."""
    import torch
    from mfgp_coverage_b200._batched import BatchedRuns
    truth_arr, prior_arr, hyp = _inputs(32, True)
    R, A, T = 6, 5, 20
    starts = np.stack([synth.agents(A, 40 + r) for r in range(R)])
    unif = np.random.default_rng(1).random((R, T, A))
    noise = np.random.default_rng(2).normal(0, 0.1, (R, T * A))
    a = BatchedRuns("todescato", truth_arr, prior_arr, hyp, A, T, starts, uniforms=unif, noise=noise).run()
    b = BatchedRuns("todescato", truth_arr, prior_arr, hyp, A, T, starts, uniforms=unif, noise=noise).run(use_graph=True)
    assert b.graph is not None
    for name in ("log_loss", "log_agent", "log_sample", "nsamples", "Ncur", "mu", "var", "pos"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert int(a.nsamples.sum()) > 0
