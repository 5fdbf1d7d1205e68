"""Load the UNMODIFIED reference from /root/reference in this container (TEST INFRASTRUCTURE; never on the GPU box).

Recipe of SURVEY.md section 8(c): `oracle/refshim/` provides import stand-ins for autograd / matplotlib / mlrose /
plotter (none of them installed, none of them arithmetic on the hot path except matplotlib's crossings test, restated
in refshim/matplotlib/path.py); `simulator.py:190` needs `dtype=object` on numpy >= 1.24, applied to the source text
in memory -- no reference source is copied into this repo.  Determinism: `random.seed(s)` and
`simulator.np.random.default_rng` replaced by a function returning one shared seeded Generator.
"""
import importlib
import os
import random
import sys
import types

import numpy as np

REF_ROOT = "/root/reference"
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "simulator.py"))


_cache = {}


def load():
    """Returns (simulator_module, gaussian_process_module) of the live reference."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise RuntimeError("/root/reference is not present (GPU box?) -- use the committed fixtures in tests/golden/")
    saved = list(sys.path)
    sys.path[:0] = [_SHIM, REF_ROOT]
    try:
        for name in ("autograd", "autograd.numpy", "matplotlib", "matplotlib.pyplot", "matplotlib.path", "mlrose",
                     "plotter", "gaussian_process"):
            sys.modules.pop(name, None)
        gp = importlib.import_module("gaussian_process")
        with open(os.path.join(REF_ROOT, "simulator.py")) as f:
            src = f.read()
        patched = src.replace("np.array(vor.regions)[", "np.array(vor.regions, dtype=object)[")
        assert patched != src
        sim = types.ModuleType("reference_simulator")
        sim.__file__ = os.path.join(REF_ROOT, "simulator.py")
        exec(compile(patched, sim.__file__, "exec"), sim.__dict__)
    finally:
        sys.path[:] = saved
    _cache["mods"] = (sim, gp)
    return sim, gp


class SharedNoise:
    """Stands in for `np.random.default_rng` inside the reference: every call returns the same seeded Generator."""

    def __init__(self, seed):
        self.gen = np.random.default_rng(seed)

    def __call__(self, *a, **k):
        return self.gen


def seeded(sim, seed):
    """Seed the two RNG streams the reference draws from (simulator.py:943 and :707/:877/:1069)."""
    random.seed(seed)
    noise = SharedNoise(seed)
    sim.np = _NumpyProxy(noise)
    return noise


class _NumpyProxy:
    """`simulator.np` with only `random.default_rng` replaced (the real numpy module is left untouched)."""

    def __init__(self, noise):
        self.random = types.SimpleNamespace(default_rng=noise)

    def __getattr__(self, name):
        return getattr(np, name)
