"""mfgp-coverage_b200 -- B200-native hot path of MSU-dcypherlab/mfgp-coverage behind the reference's Python surface.

Modules mirror the reference's file names: `gaussian_process` (SFGP / MFGP), `simulator` (coverage functions and the
lloyd / periodic / todescato / choi loops) and `runner` (batch driver, CSV layout).  The arithmetic lives in
`csrc/libmfgp_b200.so` (hand-written sm_100a CUDA, C-ABI in include/mfgp_b200.h); importing this package does not
need a GPU, running anything does -- there is no CPU fallback.

The directory name carries a hyphen (it is the name the project brief gives it); `import mfgp_coverage_b200` works
through the alias module at the repository root.
"""
__version__ = "0.1.0"
