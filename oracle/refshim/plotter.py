class Plotter:  # the reference's plotter needs matplotlib; every batch run uses plotter=None (runner.py:90)
    pass
