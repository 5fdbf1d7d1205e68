"""Generate tests/golden/*.npz from the UNMODIFIED live reference and from the reference's own logged CSVs.

Runs only in the build container (needs /root/reference); the outputs are committed and are what the GPU box uses.
    python -m oracle.make_golden
Contents (SURVEY.md section 4.1 says which logged files still pin today's code):
  inputs_australia6.npz      truth grid, 9-point lofi prior, MF/SF hyper-parameters (inputs, config c2)
  inputs_two_corners.npz     same for the two_corners data set (raw-mean convention runs)
  inputs_australia3.npz      same for australia3 (BASELINE config 1: todescato_nsf, 4 agents, Data/australia3.md:11)
  ref_gp_cases.npz           live reference: SFGP/MFGP .predict mean + diag(cov) for several model states
  ref_coverage_cases.npz     live reference: compute_loss / compute_centroids / compute_max_var + Qhull polygons
  ref_runs.npz               live reference: seeded lloyd / todescato / periodic / choi runs (loss, agent, sample logs)
  ref_train.npz              live reference: NLML of MFGP / SFGP .likelihood on Data/australia6_{lofi,hifi}_train.csv rows
                             at three hyper-parameter vectors each + finite-difference gradients of the same function
  csv_headers.json           header + first row of Data/australia6_{todescato_hmf,lloyd}_{loss,agent,sample}.csv
  logged_ex_gp.npz           Data/ex_gp.csv iteration 0 (per-point Mu, Var; raw-mean convention) + its inputs
  logged_australia6_lloyd.npz  Data/australia6_lloyd_{agent,loss}.csv sims 0,1 (loss + centroid chain, 120 it)
  logged_two_corners_hmf.npz   Data/two_corners_todescato_hmf_* sim 0 (samples, centroids, VarMax; raw means)
  logged_australia6_nsf.npz    Data/australia6_todescato_nsf_* sim 0 (SF, null prior, current convention)
  logged_australia3_nsf.npz    Data/australia3_todescato_nsf_* sim 0 (config 1; pins the SF VARIANCE path only: the run
                               predates the exp(mean) convention)
"""
import os
import random

import numpy as np
import pandas as pd

from . import reference_live as rl

D = os.path.join(rl.REF_ROOT, "Data")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _csv(name):
    return pd.read_csv(os.path.join(D, name))


def _save(name, **arrays):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.0f} KiB)")


def inputs():
    for ds in ("australia6", "two_corners", "australia3"):
        _save(f"inputs_{ds}.npz",
              truth=_csv(f"{ds}_hifi.csv").values.astype(np.float64),
              prior=_csv(f"{ds}_prior.csv").values.astype(np.float64),
              mf_hyp=_csv(f"{ds}_mf_hyp.csv").values[0].astype(np.float64),
              sf_hyp=_csv(f"{ds}_sf_hyp.csv").values[0].astype(np.float64))


def gp_cases(sim, gp):
    truth = _csv("australia6_hifi.csv")
    mf, sf = _csv("australia6_mf_hyp.csv"), _csv("australia6_sf_hyp.csv")
    prior = _csv("australia6_prior.csv")
    tarr = truth.values.astype(np.float64)
    xs = tarr[:, :2]
    rng = np.random.default_rng(42)
    out = {}
    k = 0
    for kind, nh in (("M", 0), ("M", 7), ("M", 60), ("M", 250), ("S", 0), ("S", 33), ("S", 180), ("Snull", 25)):
        idx = rng.choice(xs.shape[0], nh, replace=False)
        XH, yH = tarr[idx, :2], tarr[idx, 2:3] + 0.1 * rng.standard_normal((nh, 1))
        if kind == "M":
            m = sim.init_MFGP(mf, prior)
            m.updt_info(m.X_L, m.y_L, XH, yH)
            XL, yL, hyp = m.X_L, m.y_L, mf.values[0]
        else:
            m = sim.init_SFGP(sf, prior if kind == "S" else None)
            m.updt_info(np.vstack((m.X, XH)), np.vstack((m.y, yH)))
            XL, yL, XH, yH, hyp = np.empty((0, 2)), np.empty((0, 1)), m.X, m.y, sf.values[0]
        mu, cov = m.predict(xs)
        out.update({f"c{k}_hyp": hyp.astype(np.float64), f"c{k}_XL": XL, f"c{k}_yL": yL, f"c{k}_XH": XH, f"c{k}_yH": yH,
                    f"c{k}_mu": mu[:, 0], f"c{k}_var": np.diag(cov).copy(), f"c{k}_L": m.L})
        k += 1
    _save("ref_gp_cases.npz", ncases=np.array(k), xs=xs, **out)


def coverage_cases(sim, gp):
    truth = _csv("australia6_hifi.csv").values.astype(np.float64)
    xs = truth[:, :2]
    bbox = np.array([xs[:, 0].min(), xs[:, 0].max(), xs[:, 1].min(), xs[:, 1].max()])
    rng = np.random.default_rng(5)
    out = {}
    k = 0
    for A, on_grid in ((4, False), (8, False), (8, True), (16, True), (16, False), (64, False)):
        seeds = rng.random((A, 2))
        if on_grid:
            seeds = xs[rng.choice(xs.shape[0], A, replace=False)].copy()
        mu = rng.normal(0.3, 0.2, (xs.shape[0], 1))
        var = rng.random(xs.shape[0])
        var[rng.choice(xs.shape[0], 30)] = var.max()
        vor = sim.voronoi_bounded(seeds, bbox)
        loss = sim.compute_loss(vor, truth)
        cen = sim.compute_centroids(vor, xs, mu)
        axy, mv = sim.compute_max_var(vor, truth, np.diag(var))
        member = np.stack([sim.in_polygon(xs[:, 0], xs[:, 1], vor.vertices[c, 0], vor.vertices[c, 1])
                           for c in vor.filtered_regions])
        off = np.cumsum([0] + [len(c) for c in vor.filtered_regions]).astype(np.int32)
        poly = np.concatenate([vor.vertices[c, :] for c in vor.filtered_regions], axis=0)
        out.update({f"c{k}_seeds": seeds, f"c{k}_mu": mu[:, 0], f"c{k}_var": var, f"c{k}_loss": np.array(loss),
                    f"c{k}_cent": cen, f"c{k}_argmax_xy": axy, f"c{k}_maxvar": mv[:, 0],
                    f"c{k}_member": np.packbits(member, axis=1), f"c{k}_poly": poly, f"c{k}_off": off})
        k += 1
    _save("ref_coverage_cases.npz", ncases=np.array(k), truth=truth, **out)


def _logs_to_arrays(prefix, logs, agents):
    loss_log, agent_log, sample_log = logs
    cols = ("X", "Y", "XMax", "YMax", "VarMax", "Var0", "XCentroid", "YCentroid", "ProbExplore", "Explore", "Distance")
    T = len(loss_log)
    agent = np.array([[float(r[c]) for c in cols] for r in agent_log]).reshape(T, agents, len(cols))
    samples = np.array([[float(r["Iteration"]), float(r["Agent"]), float(r["X"]), float(r["Y"]), float(r["Sample"])]
                        for r in sample_log if r["Agent"] != "NA"]).reshape(-1, 5)
    return {f"{prefix}_loss": np.array([r["Loss"] for r in loss_log]), f"{prefix}_agent": agent,
            f"{prefix}_samples": samples, f"{prefix}_period": np.array([r["Period"] for r in loss_log])}


def ref_runs(sim, gp):
    out = {"agent_cols": np.array(["X", "Y", "XMax", "YMax", "VarMax", "Var0", "XCentroid", "YCentroid", "ProbExplore",
                                   "Explore", "Distance"])}
    for ds, cases in (("australia6", (("lloyd", "lloyd", "sf", "null", 8, 20, 11),
                                      ("todescato_hmf", "todescato", "mf", "prior", 8, 16, 12),
                                      ("todescato_nsf", "todescato", "sf", "null", 4, 16, 13),
                                      ("periodic_hsf", "periodic", "sf", "prior", 8, 14, 14),
                                      ("periodic_hmf", "periodic", "mf", "prior", 8, 14, 15),
                                      ("choi_hmf", "choi", "mf", "prior", 8, 24, 16),
                                      ("choi_nsf", "choi", "sf", "null", 4, 24, 17))),
                      ("australia3", (("todescato_nsf", "todescato", "sf", "null", 4, 30, 31),      # BASELINE config 1
                                      ("choi_nsf", "choi", "sf", "null", 4, 24, 32)))):
        truth = _csv(f"{ds}_hifi.csv")
        hyps = {"mf": _csv(f"{ds}_mf_hyp.csv"), "sf": _csv(f"{ds}_sf_hyp.csv")}
        priors = {"prior": _csv(f"{ds}_prior.csv"), "null": _csv("null_prior.csv")}
        for name, algo, h, pr, A, T, seed in cases:
            rl.seeded(sim, seed)
            pos = np.column_stack(([random.random() for _ in range(A)], [random.random() for _ in range(A)]))
            start = pos.copy()
            logs = getattr(sim, algo)(name, 0, T, A, pos, truth, 0.1, priors[pr], hyps[h], False, None, True)
            key = f"{ds}_{name}"
            out.update(_logs_to_arrays(key, logs, A))
            out[f"{key}_meta"] = np.array([A, T, seed])
            out[f"{key}_start"] = start
            print(key, "iterations", len(logs[0]), "samples", out[f"{key}_samples"].shape[0])
    _save("ref_runs.npz", **out)


def logged():
    # Data/ex_gp.csv: per-point Mu/Var of sim 0 iteration 0 (model = 25-point lofi prior only), raw-mean convention
    g = _csv("ex_gp.csv")
    g0 = g[(g.SimNum == 0) & (g.Iteration == 0)]
    hyp = pd.read_csv(os.path.join(D, "ex_hyp.csv")).iloc[0].values.astype(np.float64)
    pr = _csv("ex_prior.csv")
    _save("logged_ex_gp.npz", hyp=hyp, prior=pr[["X", "Y", "Means"]].values.astype(np.float64),
          xs=g0[["X", "Y"]].values.astype(np.float64), mu=g0.Mu.values.astype(np.float64),
          var=g0.Var.values.astype(np.float64))
    # australia6 lloyd: sims 0 and 1
    a, l = _csv("australia6_lloyd_agent.csv"), _csv("australia6_lloyd_loss.csv")
    out = {}
    for s in (0, 1):
        aa = a[a.SimNum == s].sort_values(["Iteration", "Agent"])
        T = aa.Iteration.max() + 1
        A = aa.Agent.max() + 1
        out[f"s{s}_pos"] = aa[["X", "Y"]].values.reshape(T, A, 2)
        out[f"s{s}_cent"] = aa[["XCentroid", "YCentroid"]].values.reshape(T, A, 2)
        out[f"s{s}_loss"] = l[l.SimNum == s].sort_values("Iteration").Loss.values
    _save("logged_australia6_lloyd.npz", **out)
    # replayable GP runs: logged samples + logged centroids (seeds of the next iteration) + VarMax / XMax
    for fname, stem, T in (("logged_two_corners_hmf.npz", "two_corners_todescato_hmf", 40),
                           ("logged_australia6_nsf.npz", "australia6_todescato_nsf", 30),
                           ("logged_australia3_nsf.npz", "australia3_todescato_nsf", 40)):
        a, smp = _csv(f"{stem}_agent.csv"), _csv(f"{stem}_sample.csv")
        aa = a[(a.SimNum == 0) & (a.Iteration < T)].sort_values(["Iteration", "Agent"])
        A = int(aa.Agent.max() + 1)
        ss = smp[(smp.SimNum == 0) & (smp.Iteration < T)]
        _save(fname, pos=aa[["X", "Y"]].values.reshape(T, A, 2), cent=aa[["XCentroid", "YCentroid"]].values.reshape(T, A, 2),
              varmax=aa.VarMax.values.reshape(T, A), xmax=aa.XMax.values.reshape(T, A), var0=aa.Var0.values.reshape(T, A),
              samples=ss[["Iteration", "Agent", "X", "Y", "Sample"]].values.astype(np.float64))


def train_cases(sim, gp):
    """Live reference `MFGP.likelihood` / `SFGP.likelihood` (gaussian_process.py:81-105, :344-384) on the reference's own
    training files (trainer.py:27, :66-67), at the constructor's initial hyper-parameters (trainer.py:35-36, :76-78), at the
    shipped trained ones and at a perturbed point; central finite differences of the same function for the gradient
    (the reference's autograd is not installed)."""
    lofi = np.loadtxt(os.path.join(D, "australia6_lofi_train.csv"), skiprows=1, delimiter=",")[:150]
    hifi = np.loadtxt(os.path.join(D, "australia6_hifi_train.csv"), skiprows=1, delimiter=",")[:150]
    mf = _csv("australia6_mf_hyp.csv").values[0].astype(np.float64)
    sf = _csv("australia6_sf_hyp.csv").values[0].astype(np.float64)
    X_L, y_L, X_H, y_H = lofi[:, :2], lofi[:, 2:3], hifi[:, :2], hifi[:, 2:3]
    out = {"X_L": X_L, "y_L": y_L, "X_H": X_H, "y_H": y_H}
    rng = np.random.default_rng(9)

    def fd(fn, h):
        g = np.zeros(h.size)
        for k in range(h.size):
            e = np.zeros(h.size)
            e[k] = 1e-5
            g[k] = (fn(h + e) - fn(h - e)) / 2e-5
        return g
    m = gp.MFGP(X_L, y_L, X_H, y_H, 0.5, 0.1)
    hyps = [m.hyp.copy(), mf, mf + 0.2 * rng.standard_normal(9)]
    out["mf_hyps"] = np.array(hyps)
    out["mf_nlml"] = np.array([m.likelihood(h) for h in hyps])
    out["mf_grad_fd"] = np.array([fd(m.likelihood, h) for h in hyps])
    s = gp.SFGP(X_H, y_H, 0.01)
    hyps = [s.hyp.copy(), sf, sf + 0.2 * rng.standard_normal(4)]
    out["sf_hyps"] = np.array(hyps)
    out["sf_nlml"] = np.array([s.likelihood(h) for h in hyps])
    out["sf_grad_fd"] = np.array([fd(s.likelihood, h) for h in hyps])
    _save("ref_train.npz", **out)


def csv_headers():
    """First line + first data row of the reference's logged output CSVs (runner.py:150-156 writes them with pandas'
    default index column): the column ORDER analysis.py reads; checked against what runner.run() writes."""
    import json
    out = {}
    for stem in ("australia6_todescato_hmf", "australia6_lloyd"):
        for kind in ("loss", "agent", "sample"):
            with open(os.path.join(D, f"{stem}_{kind}.csv")) as f:
                out[f"{stem}_{kind}"] = {"header": f.readline().rstrip("\n"), "first_row": f.readline().rstrip("\n")}
    path = os.path.join(OUT, "csv_headers.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(f"wrote {path}")


def main():
    os.makedirs(OUT, exist_ok=True)
    csv_headers()
    sim, gp = rl.load()
    inputs()
    logged()
    gp_cases(sim, gp)
    coverage_cases(sim, gp)
    ref_runs(sim, gp)
    train_cases(sim, gp)


if __name__ == "__main__":
    main()
