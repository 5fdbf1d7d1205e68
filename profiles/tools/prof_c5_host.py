import cProfile, pstats, io, sys, os, random, contextlib
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
from tests import synth
from mfgp_coverage_b200 import simulator as sim
sim.INCREMENTAL = True
sim.VORONOI = 'clip'
truth_arr, prior_arr = bench.c5_inputs()
def one(k):
    with contextlib.redirect_stdout(io.StringIO()):
        sim.periodic("periodic_hmf", k, 120, 8, synth.agents(8, k), truth_arr, 0.1, prior_arr, synth.MF_HYP, False, None, True, rng=random.Random(k), noise_rng=np.random.default_rng(k))
one(0)
pr = cProfile.Profile(); pr.enable(); one(1); one(2); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue())
