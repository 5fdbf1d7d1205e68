"""Import shim: the reference only uses matplotlib.path.Path.contains_points on the hot path (simulator.py:123-124)."""
