"""cProfile of the host side of the batched c5 sweep (simulator.run_batched, 512 runs x 120 iterations): where the ~0.6 s
outside the device loop goes.  usage: prof_c5_batched_host.py [runs=512]"""
import cProfile, pstats, io, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench
from tests import synth
from mfgp_coverage_b200 import simulator as sim
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 512
truth_arr, prior_arr = bench.c5_inputs()
A, T = bench.C5["agents"], bench.C5["iterations"]
def sweep(tag):
    starts = np.stack([synth.agents(A, 100_000 * tag + k) for k in range(runs)])
    rngs = [np.random.default_rng(100_000 * tag + k) for k in range(runs)]
    return sim.run_batched("periodic_hmf", list(range(runs)), T, A, starts, truth_arr, bench.C5["sigma_n"], prior_arr, synth.MF_HYP,
                           noise_rngs=rngs, use_graph=True)
sweep(1)
pr = cProfile.Profile(); pr.enable(); sweep(2); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue())
