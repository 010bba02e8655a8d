"""mgm_b200 -- B200-native MGM stereo hot path (cost volume -> multi-direction
aggregation -> WTA + sub-pixel) behind a thin C ABI (include/mgmb200.h).

This package is a ctypes binding of ``libmgmb200.so`` (hand-written sm_100a CUDA,
built in-tree by ``mgm_b200/csrc/Makefile``) plus a numpy-facing mirror of the
reference's operator interface (``compute_mgm_weights``,
``allocate_and_fill_sgm_costvolume``, ``mgm``, ``subpixel_refinement_sgm``).
There is no CPU path: if the library is missing or no CUDA device is present the
calls raise.
"""
from .api import (Context, MgmError, build_library, library_path, load_library, DISTANCES, PREFILTERS,
                  REFINEMENTS, StereoParams, exported_symbols)

__all__ = ["Context", "MgmError", "build_library", "library_path", "load_library", "DISTANCES", "PREFILTERS",
           "REFINEMENTS", "StereoParams", "exported_symbols"]
