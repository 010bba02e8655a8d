"""CPU suite, part 2: the C-ABI library loads, exports every symbol include/mgmb200.h declares, its
name tables mirror the reference's silent fallbacks, and it fails loudly without a GPU (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import mgm_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = mgm_b200.load_library()
    names = mgm_b200.exported_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
    assert lib.mgmb200_version() == 100


def test_header_cites_reference_interfaces():
    text = open(os.path.join(ROOT, "include", "mgmb200.h")).read()
    for cite in ["mgm_weights.h:63", "mgm_costvolume.h:337", "mgm_core.cc:408", "mgm_refine.h:40", "matlab/mgm_o.cc"]:
        assert cite in text


def test_name_tables():
    lib = mgm_b200.load_library()
    for i, n in enumerate(mgm_b200.DISTANCES):
        assert lib.mgmb200_distance_index(n.encode()) == i
    for i, n in enumerate(mgm_b200.PREFILTERS):
        assert lib.mgmb200_prefilter_index(n.encode()) == i
    for i, n in enumerate(mgm_b200.REFINEMENTS):
        assert lib.mgmb200_refinement_index(n.encode()) == i
    assert lib.mgmb200_distance_index(b"l1") == 0          # unknown -> 0 (mgm_costvolume.h:184-190)
    assert lib.mgmb200_prefilter_index(b"sobel_x") == 0    # the reference's own Makefile:18 typo
    assert lib.mgmb200_refinement_index(b"quartic") == 0


def test_padded_layout_helpers():
    lib = mgm_b200.load_library()
    for L in [1, 2, 31, 32, 33, 151, 256]:
        VS = lib.mgmb200_padded_labels(L)
        assert VS >= L and VS % 32 == 0 and VS - L < 32
        assert lib.mgmb200_volume_bytes(7, 5, L) == 7 * 5 * VS * 4


def test_default_params_match_cli_defaults():
    lib = mgm_b200.load_library()
    p = mgm_b200.StereoParams()
    lib.mgmb200_stereo_params_default(ctypes.byref(p))
    assert (p.dmin, p.dmax, p.NDIR, p.MGM) == (-30, 30, 4, 4)          # mgm.cc:305-307,186
    assert (p.P1, p.P2, p.aP, p.aThresh) == (8.0, 32.0, 1.0, 5.0)       # mgm.cc:308-312
    assert (p.prefilter, p.distance, p.refinement) == (b"none", b"ad", b"none")
    assert np.isinf(p.truncDist) and p.census_ncc_win == 3 and p.sgm_fix_overcount == 1


def test_no_silent_cpu_fallback():
    """Without a CUDA device the product must refuse to run (and never import the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mgm_b200.MgmError) as e:
        mgm_b200.Context()
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mgm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle|liboracle|libmgmref|#include\s+\"[^\"]*oracle", src, re.M), f
