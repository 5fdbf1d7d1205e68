"""Batched replicate runs: many independent simulations of ONE experiment stepped together on the device.

The reference runs its replicate sweeps (100+ simulations of the same algorithm on the same data, runner.py:100, :131-147)
one simulation per worker process; every simulation is a host loop that calls the GP and the coverage functions once per
iteration (simulator.py periodic :618-785, todescato :788-954, lloyd :508-616).  On small grids that loop is launch-bound on
a GPU.  `BatchedRuns` keeps the loop state of all runs on the device and advances every run by one iteration with three
kernel launches (`mfgp_batch_step`, csrc/batched.cu); the host draws the runs' random numbers up front, in the order the
reference consumes them, and reads the log buffers back once at the end.  Results per run: the reference's three lists of
log-row dicts.

Parity: samples, decisions, centroids, arg-max points and variances follow the single-run path (appended samples border the
factor instead of a refit: same results to ~1e-13).  The LOSS of an iteration in which agents sit on grid points can involve
grid points lying exactly on a bisector; the reference resolves those through Qhull's vertex rounding.  The stepper counts
them per (run, iteration); `exact_tie_loss=True` re-evaluates exactly those losses with Qhull cells on the single-run path.
"""
import ctypes

import numpy as np
import torch

from . import _coverage as cv
from . import _native as nat
from .gaussian_process import evaluate_hyp, prior_variance

ALGOS = {"lloyd": 0, "periodic": 1, "todescato": 2}
LOG_COLS = ("X", "Y", "XMax", "YMax", "VarMax", "Var0", "XCentroid", "YCentroid", "ProbExplore", "Explore", "Distance")


class BatchedRuns:
    def __init__(self, algo, truth_arr, prior_arr, hyp, agents, iterations, positions0, sigma_n=0.1, uniforms=None, noise=None,
                 noise_rngs=None, raw_means=False, device=None):
        """algo: "lloyd" | "periodic" | "todescato".  positions0[R, A, 2]: start positions of the R runs.
        uniforms[R, iterations, A]: todescato's Bernoulli draws (random.random(), simulator.py:943), in consumption order.
        noise[R, iterations * A]: the N(0, sigma_n) sample noise in consumption order -- or `noise_rngs`, one numpy Generator
        per run, from which it is drawn here (the reference draws one normal per sample, :707)."""
        from . import simulator as sim
        from ._engine import TensorAxes, detect_tensor_grid
        nat.require_cuda()
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.algo = algo
        self.code = ALGOS[algo]
        truth_arr = np.ascontiguousarray(truth_arr, dtype=np.float64)
        self.truth_arr = truth_arr
        xy = np.ascontiguousarray(truth_arr[:, :2])
        tg = detect_tensor_grid(xy)
        if tg is None:
            raise ValueError("BatchedRuns needs a tensor-product grid (every grid of the reference: distribution.py:337-339)")
        ux, uy = tg
        positions0 = np.ascontiguousarray(positions0, dtype=np.float64)
        R, A = positions0.shape[0], positions0.shape[1]
        if A != agents or A > 16:
            raise ValueError("positions0 must be [runs, agents, 2] with at most 16 agents")
        G, nx, ny, T = xy.shape[0], len(ux), len(uy), int(iterations)
        self.R, self.A, self.G, self.T = R, A, G, T
        f64 = dict(dtype=torch.float64, device=self.dev)
        i32 = dict(dtype=torch.int32, device=self.dev)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        self.xy, self.f, self.ux, self.uy = t(xy), t(truth_arr[:, 2]), t(ux), t(uy)
        self.bbox = np.array([xy[:, 0].min(), xy[:, 0].max(), xy[:, 1].min(), xy[:, 1].max()])
        # ---- the shared initial model (prior only), fitted and evaluated once by the single-run engine
        self.fidelity = "NA"
        NL = N0 = 0
        cap = 8
        self.params = None
        prob0 = 0.0
        if self.code != 0:
            hyp = np.asarray(hyp.values.tolist()[0] if hasattr(hyp, "values") else hyp, dtype=np.float64).reshape(-1)
            self.fidelity = sim._fidelity_of(hyp)
            model = sim.init_SFGP(hyp, prior_arr) if self.fidelity == "S" else sim.init_MFGP(hyp, prior_arr)
            model.raw_means = raw_means
            if self.fidelity == "S":
                model.updt_info(model.X, model.y)
                X0, y0 = np.asarray(model.X, dtype=np.float64).reshape(-1, 2), np.asarray(model.y, dtype=np.float64).reshape(-1)
            else:
                model.updt_info(model.X_L, model.y_L, model.X_H, model.y_H)
                X0, y0 = np.asarray(model.X_L, dtype=np.float64).reshape(-1, 2), np.asarray(model.y_L, dtype=np.float64).reshape(-1)
                NL = X0.shape[0]
            self.params = evaluate_hyp(hyp, raw_means)
            self.pstruct = model.engine.pstruct
            N0 = X0.shape[0]
            grid = cv.CoverageGrid(xy, truth_arr[:, 2], device=self.dev)
            mu0, var0 = model.predict_device(grid.xy, grid=grid)
            eng = model.engine
            eng.ensure_factor(need_inverse=True)
            eng.check_factor(force=True)
            cap = -(-(N0 + T * A) // 8) * 8
            k0 = prior_variance(self.params)
            if self.code == 2:
                prob0 = float(np.sqrt(float(var0.max().item()) / (k0 * A)))      # simulator.py:855-857
        else:
            self.pstruct = nat.MfgpParams(0.0, 1.0, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1e-8, 0, 0)
        self.cap, self.NL, self.N0 = cap, NL, N0
        self.max_samples = max(T * A, 1)
        z = lambda *shape, **k: torch.zeros(shape, **(k or f64))
        self.Xt, self.y, self.W, self.z = z(R, cap, 2), z(R, cap), torch.empty((R, cap, cap), **f64), z(R, cap)
        self.TxL, self.TyL, self.TxH, self.TyH = z(R, cap, nx), z(R, cap, ny), z(R, cap, nx), z(R, cap, ny)
        self.mu, self.var = z(R, G), z(R, G)
        if self.code != 0:
            if N0:
                self.Xt[:, :N0] = t(X0)
                self.y[:, :N0] = t(y0)
                self.W[:, :N0, :N0] = torch.tril(eng.W[:N0, :N0])
                self.z[:, :N0] = eng.z[:N0]
                p = self.params
                X0d = t(X0)
                for tab, u, col, l in ((self.TxH, self.ux, 0, p["l_H"]), (self.TyH, self.uy, 1, p["l_H"]),
                                       (self.TxL, self.ux, 0, p["l_L"]), (self.TyL, self.uy, 1, p["l_L"])):
                    d = u[None, :] / l - (X0d[:, col] / l)[:, None]          # the division first, as gaussian_process.py:77-78
                    tab[:, :N0] = torch.exp(-0.5 * (d * d))
            self.mu[:] = mu0
            self.var[:] = var0
        pos = t(positions0.reshape(R, A * 2))
        self.pos, self.prev, self.cen = pos.clone(), pos.clone(), pos.clone()
        self.pos_idx = torch.full((R, A), -1, dtype=torch.int64, device=self.dev)
        self.prob = torch.full((R, A), prob0, **f64)
        self.explore, self.Ncur = z(R, A, **i32), torch.full((R,), N0, **i32)
        self.knew, self.status, self.noise_used, self.nsamples = z(R, **i32), z(R, **i32), z(R, **i32), z(R, **i32)
        self.ties = z(R, T, **i32)
        if noise is None:
            gens = noise_rngs if noise_rngs is not None else [np.random.default_rng() for _ in range(R)]
            noise = np.stack([g.normal(loc=0, scale=sigma_n, size=self.max_samples) for g in gens]) if self.code else np.zeros((R, 1))
        self.noise = t(np.asarray(noise, dtype=np.float64).reshape(R, -1))
        if self.noise.shape[1] < self.max_samples and self.code:
            raise ValueError("noise must hold iterations * agents draws per run")
        self.max_samples = int(self.noise.shape[1]) if self.code else 1
        if self.code == 2:
            if uniforms is None:
                raise ValueError("todescato needs the runs' uniform draws (random.random(), simulator.py:943)")
            self.unif = t(np.asarray(uniforms, dtype=np.float64).reshape(R, T, A))
        else:
            self.unif = z(1)
        self.log_loss, self.log_agent = z(R, T), z(R, T, A, len(LOG_COLS))
        self.log_sample = z(R, self.max_samples, 5)
        b = nat.MfgpBatch()
        b.runs, b.G, b.nx, b.ny, b.A, b.NL, b.cap, b.algo, b.iterations, b.max_samples = R, G, nx, ny, A, NL, cap, self.code, T, \
            self.max_samples
        b.xmin, b.xmax, b.ymin, b.ymax = (float(v) for v in self.bbox)
        b.eps, b.tie_tol, b.amax_rel = cv.EPS, cv.TIE_TOL, cv.AMAX_REL
        for name in ("xy", "f", "ux", "uy", "Xt", "y", "W", "z", "TxL", "TyL", "TxH", "TyH", "mu", "var", "pos", "prev", "cen",
                     "pos_idx", "prob", "explore", "Ncur", "knew", "status", "noise_used", "nsamples", "ties", "noise", "unif",
                     "log_loss", "log_agent", "log_sample"):
            setattr(b, name, getattr(self, name).data_ptr())
        self._b = b
        self.graph = None

    def step(self, it):
        nat.check(nat.lib().mfgp_batch_step(ctypes.byref(self._b), ctypes.byref(self.pstruct), int(it), nat.stream_ptr()),
                  "mfgp_batch_step")

    def capture(self):
        """Record the whole run -- every iteration's three launches -- into ONE CUDA graph.  Nothing in the loop depends on
        the host (sample counts, explore decisions and the growing training-set sizes live in device memory), so the graph
        is valid for any state the buffers hold when it is replayed."""
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.graph(graph, stream=side):
            for it in range(self.T):
                self.step(it)
        self.graph = graph
        return self

    def run(self, use_graph=False):
        """All iterations, no host round trip in between; returns when the device is done.  `use_graph`: replay the run
        from one CUDA graph (captured on first use) instead of 3 * iterations launches."""
        if use_graph:
            if self.graph is None:
                self.capture()
            self.graph.replay()
        else:
            for it in range(self.T):
                self.step(it)
        torch.cuda.current_stream(self.dev).synchronize()
        st = self.status.cpu().numpy()
        bad = np.nonzero(st)[0]
        if bad.size:
            r, code = int(bad[0]), int(st[bad[0]])
            if code > 0:
                raise np.linalg.LinAlgError(f"Matrix is not positive definite (pivot {code - 1}) in run {r}")
            if code == -4:
                raise ValueError("zero-size array to reduction operation maximum which has no identity")      # np.amax([])
            raise RuntimeError(f"mfgp_batch_step: run {r} stopped with status {code}")
        return self

    def logs(self, sim_nums=None, exact_tie_loss=False):
        """The reference's (loss_log, agent_log, sample_log) row lists, one triple per run."""
        R, T, A = self.R, self.T, self.A
        loss = self.log_loss.cpu().numpy().copy()
        agent = self.log_agent.cpu().numpy()
        samples = self.log_sample.cpu().numpy()
        ns = self.nsamples.cpu().numpy()
        ties = self.ties.cpu().numpy()
        if exact_tie_loss and ties.any():
            from . import simulator as sim
            for r, it in zip(*np.nonzero(ties)):
                vor = cv.BoundedVoronoi(agent[r, it, :, :2], self.bbox)
                loss[r, it] = sim.compute_loss(vor, self.truth_arr)
        self.tie_iterations = ties > 0
        sim_nums = list(range(R)) if sim_nums is None else list(sim_nums)
        fid = self.fidelity
        out = []
        for r in range(R):
            s = sim_nums[r]
            loss_log = [{"SimNum": s, "Iteration": it, "Period": 0, "Fidelity": fid, "Loss": loss[r, it]} for it in range(T)]
            rows = agent[r].tolist()
            var0 = 0 if self.code == 0 else None
            agent_log = [{"SimNum": s, "Iteration": it, "Period": 0, "Fidelity": fid, "Agent": i,
                          "X": v[0], "Y": v[1], "XMax": v[2], "YMax": v[3], "VarMax": v[4], "Var0": v[5] if var0 is None else 0,
                          "XCentroid": v[6], "YCentroid": v[7], "ProbExplore": v[8], "Explore": v[9], "Distance": v[10]}
                         for it in range(T) for i, v in enumerate(rows[it])]
            if self.code == 0:
                sample_log = [{"SimNum": s, "Iteration": it, "Period": 0, "Fidelity": fid, "Agent": "NA", "X": "NA", "Y": "NA",
                               "Sample": "NA"} for it in range(T)]
            else:
                sample_log = [{"SimNum": s, "Iteration": int(v[0]), "Period": 0, "Fidelity": fid, "Agent": v[1], "X": v[2], "Y": v[3],
                               "Sample": v[4]} for v in samples[r, :ns[r]].tolist()]
            out.append((loss_log, agent_log, sample_log))
        return out

    def final_positions(self):
        return self.pos.cpu().numpy().reshape(self.R, self.A, 2)
