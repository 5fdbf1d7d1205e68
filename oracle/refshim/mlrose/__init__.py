"""Deterministic stand-in for mlrose (not installed; TSP tour order is out of scope): identity tour."""


class TSPOpt:
    def __init__(self, length, coords=None, maximize=False):
        self.length = length
        self.coords = coords


def genetic_alg(problem, mutation_prob=0.2, max_attempts=100, random_state=None):
    return list(range(problem.length)), 0.0
