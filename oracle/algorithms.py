"""Oracle: the four control loops around the hot path (numpy restatement; TEST INFRASTRUCTURE).

Follows /root/reference/simulator.py: lloyd :508-616, periodic :618-785, todescato :788-954, choi :957-1161 and the
decision rules :457-500.  Array in / list-of-dict-rows out (same row schemas as the reference's logs, including the
`YMax` column that logs positions[i,1], :596,:754,:924,:1116).  Randomness is injected: `py_random` stands for the
`random` module (Bernoulli explore draws, :943) and `noise_rng` for the per-sample `np.random.default_rng()` (:707,
:877,:1069).  The Choi TSP tour (mlrose GA, :415-454: unpinned third-party routine) comes from the deterministic
planner of oracle/tsp.py, as in oracle/refshim/mlrose and in the product (csrc/tsp.cu).
"""
import numpy as np

from . import coverage as cov
from . import gp as ogp
from . import tsp


def _fidelity(hyp):
    n = np.asarray(hyp).reshape(-1).size
    if n == 4:
        return "S"
    if n == 9:
        return "M"
    raise TypeError("Hyperparameters must be of length 4 (single-fidelity) or 9 (multi-fidelity)")


def _agent_rows(sim_num, iteration, period, fidelity, positions, argmax_var_t, max_var_t, max_var_0, centroids_t,
                prob_explore_t, explore_t, distance):
    rows = []
    for i in range(positions.shape[0]):
        rows.append({"SimNum": sim_num, "Iteration": iteration, "Period": period, "Fidelity": fidelity, "Agent": i,
                     "X": positions[i, 0], "Y": positions[i, 1], "XMax": argmax_var_t[i, 0], "YMax": positions[i, 1],
                     "VarMax": max_var_t[i, 0], "Var0": max_var_0,
                     "XCentroid": centroids_t[i, 0], "YCentroid": centroids_t[i, 1],
                     "ProbExplore": prob_explore_t[i, 0], "Explore": explore_t[i, 0], "Distance": distance[i, 0]})
    return rows


def lloyd(sim_num, iterations, agents, positions, truth_arr):
    loss_log, agent_log, sample_log = [], [], []
    bbox = cov.bounding_box_of(truth_arr[:, :2])
    zeros = np.zeros((agents, 1))
    prev_positions = np.copy(positions)
    centroids_t = np.copy(positions)
    for iteration in range(iterations):
        distance = np.sqrt(np.sum((positions - prev_positions) ** 2, axis=1)).reshape(-1, 1)
        loss_t = cov.compute_loss(cov.voronoi_bounded(positions, bbox), truth_arr)
        lloyd_vor = cov.voronoi_bounded(centroids_t, bbox)
        centroids_t = cov.compute_centroids(lloyd_vor, truth_arr[:, [0, 1]], truth_arr[:, [2]])
        loss_log.append({"SimNum": sim_num, "Iteration": iteration, "Period": 0, "Fidelity": "NA", "Loss": loss_t})
        sample_log.append({"SimNum": sim_num, "Iteration": iteration, "Period": 0, "Fidelity": "NA",
                           "Agent": "NA", "X": "NA", "Y": "NA", "Sample": "NA"})
        agent_log.extend(_agent_rows(sim_num, iteration, 0, "NA", positions, zeros, zeros, 0, centroids_t,
                                     zeros, zeros, distance))
        prev_positions = np.copy(positions)
        positions = np.copy(centroids_t)
    return loss_log, agent_log, sample_log


def _take_samples(positions, explore_t, truth_arr, sigma_n, noise_rng):
    x_new, y_new, id_new = np.empty([0, 2]), np.empty([0, 1]), np.empty([0, 1])
    for i in range(positions.shape[0]):
        if explore_t[i] == 1:
            x_sample = positions[i, :]
            sel = np.logical_and(truth_arr[:, 0] == x_sample[0], truth_arr[:, 1] == x_sample[1])
            y_sample = truth_arr[sel, 2] + noise_rng.normal(loc=0, scale=sigma_n)
            x_new = np.vstack((x_new, x_sample))
            y_new = np.vstack((y_new, y_sample))
            id_new = np.vstack((id_new, i))
    return x_new, y_new, id_new


def _iteration_body(model, positions, centroids_t, truth_arr, x_star, bbox):
    mu, var = model.predict(x_star)
    loss_t = cov.compute_loss(cov.voronoi_bounded(positions, bbox), truth_arr)
    lloyd_vor = cov.voronoi_bounded(centroids_t, bbox)
    centroids_t = cov.compute_centroids(lloyd_vor, x_star, mu.reshape(-1, 1))
    argmax_var_t, max_var_t, _ = cov.compute_max_var(lloyd_vor, truth_arr, var)
    return loss_t, centroids_t, argmax_var_t, max_var_t


def _init_gp(hyp, prior_arr, truth_arr, raw_means):
    p = ogp.GPParams.from_hyp(hyp, raw_means=raw_means)
    x_star = truth_arr[:, [0, 1]]
    max_var_0 = p.k0       # amax of the empty model's covariance == k(0)  (simulator.py:671-672, :841-842)
    model = ogp.Model.from_prior(p, prior_arr)
    model.updt_info()
    return p, x_star, max_var_0, model


def _sample_rows(sim_num, iteration, period, fidelity, id_new, x_new, y_new):
    return [{"SimNum": sim_num, "Iteration": iteration, "Period": period, "Fidelity": fidelity,
             "Agent": id_new[i, 0], "X": x_new[i, 0], "Y": x_new[i, 1], "Sample": y_new[i, 0]}
            for i in range(id_new.size)]


def _explore_loop(kind, sim_num, iterations, agents, positions, truth_arr, sigma_n, prior_arr, hyp, py_random,
                  noise_rng, raw_means):
    fidelity = _fidelity(hyp)
    loss_log, agent_log, sample_log = [], [], []
    p, x_star, max_var_0, model = _init_gp(hyp, prior_arr, truth_arr, raw_means)
    bbox = cov.bounding_box_of(x_star)
    mu, var = model.predict(x_star)
    max_var_t = np.amax(var) * np.ones((agents, 1))
    if kind == "todescato":
        prob_explore_t = np.sqrt(max_var_t / (max_var_0 * agents))
    else:
        prob_explore_t = np.zeros((agents, 1))
    explore_t = np.zeros((agents, 1))
    prev_positions = np.copy(positions)
    centroids_t = np.copy(positions)
    for iteration in range(iterations):
        x_new, y_new, id_new = _take_samples(positions, explore_t, truth_arr, sigma_n, noise_rng)
        distance = np.sqrt(np.sum((positions - prev_positions) ** 2, axis=1)).reshape(-1, 1)
        model.append(x_new, y_new)
        loss_t, centroids_t, argmax_var_t, max_var_t = _iteration_body(model, positions, centroids_t, truth_arr,
                                                                        x_star, bbox)
        loss_log.append({"SimNum": sim_num, "Iteration": iteration, "Period": 0, "Fidelity": fidelity, "Loss": loss_t})
        agent_log.extend(_agent_rows(sim_num, iteration, 0, fidelity, positions, argmax_var_t, max_var_t, max_var_0,
                                     centroids_t, prob_explore_t, explore_t, distance))
        sample_log.extend(_sample_rows(sim_num, iteration, 0, fidelity, id_new, x_new, y_new))
        if kind == "todescato":
            prob_explore_t = np.sqrt(max_var_t / (max_var_0 * agents))
            explore_t = np.array([int(py_random.random() < c) for c in prob_explore_t]).reshape(-1, 1)
        else:
            b = (iteration // 5) % 2 == 0
            prob_explore_t = np.array([int(b) for _ in range(agents)]).reshape(-1, 1)
            explore_t = np.array([int(b) for _ in range(agents)]).reshape(-1, 1)
        prev_positions = np.copy(positions)
        for i in range(agents):
            positions[i, :] = argmax_var_t[i, :] if explore_t[i, 0] else centroids_t[i, :]
    return loss_log, agent_log, sample_log


def todescato(sim_num, iterations, agents, positions, truth_arr, sigma_n, prior_arr, hyp, py_random, noise_rng,
              raw_means=False):
    return _explore_loop("todescato", sim_num, iterations, agents, positions, truth_arr, sigma_n, prior_arr, hyp,
                         py_random, noise_rng, raw_means)


def periodic(sim_num, iterations, agents, positions, truth_arr, sigma_n, prior_arr, hyp, py_random, noise_rng,
             raw_means=False):
    return _explore_loop("periodic", sim_num, iterations, agents, positions, truth_arr, sigma_n, prior_arr, hyp,
                         py_random, noise_rng, raw_means)


def choi(sim_num, iterations, agents, positions, truth_arr, sigma_n, prior_arr, hyp, py_random, noise_rng,
         raw_means=False, fast_planner=True):
    fidelity = _fidelity(hyp)
    loss_log, agent_log, sample_log = [], [], []
    p, x_star, max_var_0, model = _init_gp(hyp, prior_arr, truth_arr, raw_means)
    bbox = cov.bounding_box_of(x_star)
    threshold = max_var_0
    iteration, period = 0, 0
    centroids_t = np.copy(positions)
    prev_positions = np.copy(positions)
    prob_explore_t = np.zeros((agents, 1))
    explore_t = np.zeros((agents, 1))
    planner = cov.compute_sample_points_fast if fast_planner else cov.compute_sample_points
    while iteration < iterations:
        threshold = 0.82 * threshold
        sample_vor = cov.voronoi_bounded(centroids_t, bbox)
        sample_points, _ = planner(model, x_star, threshold)
        tours = tsp.compute_sample_tsp(cov.compute_sample_clusters(sample_vor, sample_points))
        for _step in range(8 * 2 ** period):
            x_new, y_new, id_new = _take_samples(positions, explore_t, truth_arr, sigma_n, noise_rng)
            distance = np.sqrt(np.sum((positions - prev_positions) ** 2, axis=1)).reshape(-1, 1)
            model.append(x_new, y_new)
            loss_t, centroids_t, argmax_var_t, max_var_t = _iteration_body(model, positions, centroids_t, truth_arr,
                                                                            x_star, bbox)
            loss_log.append({"SimNum": sim_num, "Iteration": iteration, "Period": period, "Fidelity": fidelity,
                             "Loss": loss_t})
            agent_log.extend(_agent_rows(sim_num, iteration, period, fidelity, positions, argmax_var_t, max_var_t,
                                         max_var_0, centroids_t, prob_explore_t, explore_t, distance))
            sample_log.extend(_sample_rows(sim_num, iteration, period, fidelity, id_new, x_new, y_new))
            for i in range(agents):
                flag = 1 if tours[i].shape[0] > 0 else 0
                prob_explore_t[i] = flag
                explore_t[i] = flag
            prev_positions = np.copy(positions)
            for i in range(agents):
                if explore_t[i, 0]:
                    positions[i, :] = tours[i][0, :]
                    tours[i] = np.delete(tours[i], 0, axis=0)
                else:
                    positions[i, :] = centroids_t[i, :]
            iteration += 1
        period += 1
    return loss_log, agent_log, sample_log
