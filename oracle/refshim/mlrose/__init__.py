"""Stand-in for mlrose (not installed, unpinned): the LIVE reference's `genetic_alg` call
(/root/reference/simulator.py:435-438) is routed to the deterministic planner of oracle/tsp.py -- the same objective
(closed-tour length), the same tours as the product's device planner, so seeded reference Choi runs are comparable."""
import importlib.util
import os

_spec = importlib.util.spec_from_file_location(
    "_oracle_tsp", os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tsp.py"))
_tsp = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_tsp)


class TSPOpt:
    def __init__(self, length, coords=None, maximize=False):
        self.length = length
        self.coords = coords


def genetic_alg(problem, mutation_prob=0.2, max_attempts=100, random_state=None):
    tour = _tsp.plan_tour(problem.coords)
    return tour, _tsp.tour_length(problem.coords, tour)
