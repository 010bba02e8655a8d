"""CPU suite, part 1: the oracle port (oracle/mgm_oracle.c) against
 (a) the golden vectors generated from the unmodified reference (tests/golden/, always), and
 (b) the reference itself, live, when oracle/_ref/libmgmref.so is present (build container, GPU box).
Bit-exact everywhere: the port restates the same IEEE operations in the same order."""
import itertools
import os

import numpy as np
import pytest

import oracle as O
from tests.conftest import synth_pair, synth_volume, synth_weights
from tests.golden_util import golden_files, load_golden

needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("path", golden_files("pipeline"))
def test_port_matches_golden_pipeline(path):
    g = load_golden(path)
    r = O.orc_pipeline(g["u"], g["v"], **g["params"])
    assert same(r["out"], g["out"]), g["name"]
    assert same(r["outcost"], g["outcost"]), g["name"]
    cc = r["cc"]
    assert int(np.isinf(cc).sum()) == int(g["cc_ninf"][0])
    assert np.nansum(np.where(np.isfinite(cc), cc, 0), dtype=np.float64) == g["cc_sum"][0]
    assert int((r["w"] != 1).sum()) == int(g["w_not_one"][0])


@pytest.mark.parametrize("path", golden_files("volume"))
def test_port_matches_golden_volume(path):
    g = load_golden(path)
    p = g["params"]
    r = O.orc_mgm(g["cc"], g["w"], p["dmin"], p["P1"], p["P2"], p["NDIR"], p["K"], p["felz"], p["fix"])
    assert same(r["out"], g["out"]) and same(r["outcost"], g["outcost"])
    assert same(r["S"][r["S"].shape[0] // 2], g["S_row"])
    assert np.sum(np.where(np.isfinite(r["S"]), r["S"], 0), dtype=np.float64) == g["S_sum"][0]


def test_scan_space_predecessors():
    """every sweep reads the same four scan-space predecessors; sweeps 0-3 and 4-7 differ
    only in the order."""
    axis = [(-1, 0), (0, -1), (-1, -1), (1, -1)]
    diag = [(1, -1), (-1, -1), (0, -1), (-1, 0)]
    for p in range(8):
        assert O.orc_scan_preds(p) == (axis if p < 4 else diag)


def test_name_tables_fall_back_to_zero():
    lib = O.orc_lib()
    assert lib.orc_distance_index(b"census") == 2 and lib.orc_distance_index(b"nonsense") == 0
    assert lib.orc_prefilter_index(b"sobelx") == 2 and lib.orc_prefilter_index(b"sobel_x") == 0   # Makefile:18
    assert lib.orc_refinement_index(b"parabolaOCV") == 4 and lib.orc_refinement_index(b"") == 0


@needs_ref
def test_ref_flavours_agree():
    """serial, OpenMP and DVEC_ALLOCATION_HACK builds of the reference are bit-identical"""
    cc = synth_volume(33, 21, 11, seed=4, real=True)
    w = synth_weights(33, 21, seed=4)
    base = O.ref_mgm(cc, w, -10, 8, 32, 8, 4)
    for fl in ["_serial", "_flat"]:
        r = O.ref_mgm(cc, w, -10, 8, 32, 8, 4, flavour=fl)
        assert same(r["S"], base["S"]) and same(r["out"], base["out"])


@needs_ref
@pytest.mark.parametrize("dist,win", [("ad", 3), ("sd", 3), ("census", 3), ("census", 5), ("census", 7), ("ncc", 3),
                                      ("ncc", 5), ("btad", 3), ("btsd", 3)])
@pytest.mark.parametrize("nch", [1, 3])
def test_port_costvolume_vs_ref(dist, win, nch):
    u, v = synth_pair(41, 27, 14, seed=win + nch, nch=nch)
    u = u + np.float32(0.37)   # non-integer values exercise the float paths (NCC, BT halves)
    for trunc in [np.inf, 17.5]:
        a = O.orc_costvolume(u, v, -13, 3, "none", dist, trunc, win)
        b = O.ref_costvolume(u, v, -13, 3, "none", dist, trunc, win)
        assert same(a, b), (dist, win, nch, trunc)
    # different image sizes for u and v (the R->L run of the CLI uses the same sizes, the API allows any)
    a = O.orc_costvolume(u, v[:, :, :30], -13, 3, "none", dist, np.inf, win)
    b = O.ref_costvolume(u, v[:, :, :30], -13, 3, "none", dist, np.inf, win)
    assert same(a, b)


@needs_ref
def test_port_sobelx_and_weights_vs_ref():
    u, v = synth_pair(37, 23, 12, seed=9, nch=3)
    assert same(O.orc_costvolume(u, v, -11, 2, "sobelx", "ad", 50.0, 3), O.ref_costvolume(u, v, -11, 2, "sobelx", "ad", 50.0, 3))
    for aP, aT in [(4.0, 5.0), (0.5, 30.0), (1.0, 5.0)]:
        assert same(O.orc_weights(u, aP, aT), O.ref_weights(u, aP, aT))


@needs_ref
@pytest.mark.parametrize("nch", [1, 3])
def test_port_gblur_vs_ref(nch):
    """-p gblur: gblur_truncated(sigma=1) of both images (img_tools.h:143-180), then any non-census distance"""
    u, v = synth_pair(43, 21, 12, seed=11 + nch, nch=nch)
    u = u + np.float32(0.41)
    for dist, trunc in [("ad", np.inf), ("sd", 400.0), ("btad", np.inf), ("ncc", np.inf)]:
        assert same(O.orc_costvolume(u, v, -11, 2, "gblur", dist, trunc, 3), O.ref_costvolume(u, v, -11, 2, "gblur", dist, trunc, 3)), dist
    # census forces the census prefilter whatever -p says (mgm_costvolume.h:358-362)
    assert same(O.orc_costvolume(u, v, -11, 2, "gblur", "census", np.inf, 3), O.ref_costvolume(u, v, -11, 2, "gblur", "census", np.inf, 3))


@needs_ref
@pytest.mark.parametrize("K", [1, 2, 3, 4])
@pytest.mark.parametrize("felz", [0, 1])
@pytest.mark.parametrize("weighted", [0, 1])
def test_port_mgm_vs_ref(K, felz, weighted):
    nx, ny, L = 29, 19, 10
    for real, NDIR, fix, (P1, P2) in itertools.product([0, 1], [1, 3, 8], [0, 1], [(8, 32), (2, 20000), (1.3, 7.7)]):
        cc = synth_volume(nx, ny, L, seed=NDIR, real=bool(real))
        w = synth_weights(nx, ny, seed=K) if weighted else None
        a = O.orc_mgm(cc, w, -(L - 1), P1, P2, NDIR, K, felz, fix)
        b = O.ref_mgm(cc, w, -(L - 1), P1, P2, NDIR, K, felz, fix)
        assert same(a["S"], b["S"]) and same(a["out"], b["out"]) and same(a["outcost"], b["outcost"]), \
            (real, NDIR, fix, P1, P2)


@needs_ref
@pytest.mark.parametrize("method", O.REFINEMENTS + ["bogus"])
def test_port_refine_vs_ref(method):
    cc = synth_volume(41, 25, 13, seed=2, real=True)
    r = O.ref_mgm(cc, None, -12, 8, 32, 8, 2)
    a = O.orc_refine(r["S"], -12, r["out"], r["outcost"], method)
    b = O.ref_refine(r["S"], -12, r["out"], r["outcost"], method)
    assert same(a[0], b[0]) and same(a[1], b[1])


@needs_ref
def test_port_pipeline_vs_ref():
    u, v = synth_pair(61, 37, 20, seed=5)
    for kw in [dict(distance="census", win=5, NDIR=8, K=2, refinement="vfit"),
               dict(distance="census", win=3, NDIR=8, K=3, felz=1, P1=2.0, P2=20000.0, refinement="vfit"),
               dict(distance="ad", NDIR=4, K=4, aP=4.0, aThresh=6.0, refinement="cubic")]:
        a = O.orc_pipeline(u, v, -19, 0, **kw)
        b = O.ref_pipeline(u, v, -19, 0, **kw)
        assert same(a["out"], b["out"]) and same(a["outcost"], b["outcost"]), kw


def _ragged(nx, ny, emin, emax, seed):
    rng = np.random.default_rng(seed)
    lo = rng.integers(emin, emin + 6, (ny, nx)).astype(np.float32)
    hi = (lo + rng.integers(3, 9, (ny, nx))).clip(max=emax).astype(np.float32)
    return lo, hi


@needs_ref
def test_port_ranges_vs_ref():
    """per-pixel disparity ranges: cost volume, every message update variant, WTA restricted to a
    second set of ranges (shrunk, and grown beyond the cost ranges as update_dmin_dmax does), refinement"""
    nx, ny, emin, emax = 31, 19, -12, 3
    u, v = synth_pair(nx, ny, 14, seed=3, nch=1)
    lo, hi = _ragged(nx, ny, emin, emax, 0)
    for dist in ["ad", "census", "ncc", "btsd"]:
        assert same(O.orc_costvolume_ranges(u, v, lo, hi, emin, emax, "none", dist, np.inf, 3),
                    O.ref_costvolume_ranges(u, v, lo, hi, emin, emax, "none", dist, np.inf, 3)), dist
    cc = O.ref_costvolume_ranges(u, v, lo, hi, emin, emax, "none", "ad", np.inf, 3)
    srs = [(lo, hi), (lo + 1, np.maximum(hi - 1, lo + 1)), (np.maximum(lo - 2, emin), np.minimum(hi + 2, emax))]
    for K, felz, weighted, fix in itertools.product([1, 2, 3, 4], [0, 1], [0, 1], [0, 1]):
        w = synth_weights(nx, ny, seed=K) if weighted else None
        P1, P2 = (8, 32) if not felz else (2, 20000)
        for slo, shi in srs:
            a = O.orc_mgm_ranges(cc, lo, hi, w, emin, slo, shi, P1, P2, 8, K, felz, fix)
            b = O.ref_mgm_ranges(cc, lo, hi, w, emin, slo, shi, P1, P2, 8, K, felz, fix)
            ok = np.isfinite(b["outcost"])   # the reference leaves the label uninitialised when nothing is finite
            assert same(a["S"], b["S"]) and same(a["outcost"], b["outcost"]), (K, felz, weighted, fix)
            assert same(a["out"][ok], b["out"][ok]), (K, felz, weighted, fix)
            ra = O.orc_refine_ranges(a["S"], slo, shi, emin, b["out"], b["outcost"], "vfit")
            rb = O.ref_refine_ranges(b["S"], slo, shi, emin, b["out"], b["outcost"], "vfit")
            assert same(ra[0][ok], rb[0][ok]) and same(ra[1][ok], rb[1][ok])


def _disparity_maps(nx, ny, seed):
    """a pair of plausible L->R / R->L maps with sub-pixel parts, NaNs, infinities and outliers"""
    rng = np.random.default_rng(seed)
    base = -np.round(6 + 4 * np.sin(0.11 * np.arange(nx))[None, :] * np.cos(0.07 * np.arange(ny))[:, None])
    dl = (base + rng.integers(-1, 2, (ny, nx)) * (rng.random((ny, nx)) < 0.2)).astype(np.float32)
    dl += (rng.random((ny, nx)) < 0.3) * rng.uniform(-0.5, 0.5, (ny, nx)).astype(np.float32)
    dr = (-base + rng.integers(-2, 3, (ny, nx)) * (rng.random((ny, nx)) < 0.2)).astype(np.float32)
    for d in (dl, dr):
        d[rng.random((ny, nx)) < 0.05] = np.nan
        d[rng.random((ny, nx)) < 0.01] = np.inf
        d[rng.random((ny, nx)) < 0.01] = 1e9
    return dl, dr


@pytest.mark.skipif(O.refcli_lib() is None, reason="oracle/_ref/libmgmref_cli.so not built")
def test_port_postprocessing_vs_ref():
    """leftright_test, median_filter, update_dmin_dmax of the port against the reference's own functions"""
    for seed, (nx, ny) in enumerate([(37, 23), (64, 5), (3, 3), (1, 7)]):
        dl, dr = _disparity_maps(nx, ny, seed)
        for tau in (1.0, 0.0, 2.5):
            assert same(O.orc_leftright(dl, dr, tau), O.ref_leftright(dl, dr, tau))
            assert same(O.orc_leftright(dr, dl, tau), O.ref_leftright(dr, dl, tau))
        for r in (1, 2, 3):
            assert same(O.orc_median(dl, r), O.ref_median(dl, r))
        lo = np.full((ny, nx), -12, np.float32)
        hi = np.full((ny, nx), 3, np.float32)
        for off in (dl, np.full((ny, nx), np.nan, np.float32)):
            a, b = O.orc_update_range(off, lo, hi), O.ref_update_range(off, lo, hi)
            assert same(a[0], b[0]) and same(a[1], b[1]) and a[2] == b[2]
        a, b = O.orc_update_range(dl, lo, hi, slack=-2, radius=1), O.ref_update_range(dl, lo, hi, slack=-2, radius=1)
        assert same(a[0], b[0]) and same(a[1], b[1]) and a[2] == b[2]


REF_MGM = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "mgm")


def _pnm(path, a):
    a = np.clip(a, 0, 255).astype(np.uint8)
    if a.shape[0] == 1:
        open(path, "wb").write(b"P5\n%d %d\n255\n" % (a.shape[2], a.shape[1]) + a[0].tobytes())
    else:
        open(path, "wb").write(b"P6\n%d %d\n255\n" % (a.shape[2], a.shape[1]) + np.transpose(a, (1, 2, 0)).tobytes())


@pytest.mark.skipif(not os.path.exists(REF_MGM), reason="oracle/_ref/mgm not built")
@pytest.mark.parametrize("nch,median,refinement", [(1, 1, "vfit"), (3, 0, "none")])
def test_port_cli_flow_vs_reference_cli(tmp_path, nch, median, refinement):
    """The oracle's composition of the default command-line flow (both directions, median, left-right tests,
    back-projection: the checker of mgmb200_stereo_lr) against the reference BINARY on the same files."""
    import subprocess
    u, v = synth_pair(70, 40, 14, seed=6, nch=nch)
    u, v = np.round(np.clip(u, 0, 255)), np.round(np.clip(v, 0, 255))   # what the 8-bit files hold
    if u.ndim == 2:
        u, v = u[None], v[None]
    _pnm(str(tmp_path / "u.pnm"), u)
    _pnm(str(tmp_path / "v.pnm"), v)
    names = [str(tmp_path / ("ref_%s.npy" % k)) for k in ("disp", "cost", "back")]
    env = dict(os.environ, TSGM="3", MEDIAN=str(median), USE_TRUNCATED_LINEAR_POTENTIALS="1", CENSUS_NCC_WIN="3")
    r = subprocess.run([REF_MGM, "-r", "-13", "-R", "2", "-t", "census", "-O", "8", "-P1", "2", "-P2", "20000", "-s", refinement,
                        "-l", str(tmp_path / "ref_nolr.npy"), str(tmp_path / "u.pnm"), str(tmp_path / "v.pnm")] + names,
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    disp, cost, back = [np.load(n) for n in names]
    nolr = np.squeeze(np.load(str(tmp_path / "ref_nolr.npy")))
    disp, cost = np.squeeze(disp), np.squeeze(cost)
    back = np.moveaxis(back, -1, 0) if back.ndim == 3 else back[None]
    kw = dict(P1=2.0, P2=20000.0, NDIR=8, K=3, felz=1, distance="census", win=3, refinement=refinement)
    uu, vv = u.astype(np.float32), v.astype(np.float32)
    L = O.orc_pipeline(uu, vv, -13, 2, **kw)
    R = O.orc_pipeline(vv, uu, -2, 13, **kw)
    offL = O.orc_median(L["out"], median) if median else L["out"]
    offR = O.orc_median(R["out"], median) if median else R["out"]
    assert same(nolr, offL)
    out = O.orc_leftright(offL, offR, 1.0)
    assert same(disp, out) and same(cost, L["outcost"])
    assert same(back, O.orc_backproject(out, uu, vv))


@pytest.mark.skipif(not os.path.exists(REF_MGM), reason="oracle/_ref/mgm not built")
@pytest.mark.parametrize("iters,ranges,felz,K", [(2, True, 0, 2), (3, False, 1, 3), (1, True, 1, 2), (2, True, 1, 3)])
def test_port_ranges_iterations_vs_reference_cli(tmp_path, iters, ranges, felz, K):
    """The oracle's composition of mgm.cc:372-395 (cost volume over the range images once, then per TSGM_ITER
    iteration mgm + refinement over the current ranges and update_dmin_dmax + non-finite repair: the checker of
    mgmb200_stereo_ranges) against the reference BINARY with -m/-M files and TSGM_ITER."""
    import subprocess
    nx, ny = 120, 64
    u, v = synth_pair(nx, ny, 24, seed=4, nch=1)
    u, v = np.round(np.clip(u, 0, 255)).reshape(1, ny, nx), np.round(np.clip(v, 0, 255)).reshape(1, ny, nx)
    _pnm(str(tmp_path / "u.pnm"), u)
    _pnm(str(tmp_path / "v.pnm"), v)
    # -truncDist keeps every label finite: a pixel whose range holds no finite label has an UNINITIALISED disparity in
    # the reference (mgm_core.cc:594), which update_dmin_dmax would then spread to its neighbours' ranges
    args = ["-r", "-23", "-R", "4", "-t", "ad", "-O", "8", "-s", "vfit", "-truncDist", "40"] + (["-P1", "2", "-P2", "20000"] if felz else [])
    if ranges:
        rng = np.random.default_rng(7)
        lo = (-14 + rng.integers(-6, 3, (ny, nx))).astype(np.float32)
        hi = (lo + rng.integers(0, 14, (ny, nx))).astype(np.float32)
        np.save(str(tmp_path / "dmin.npy"), lo)
        np.save(str(tmp_path / "dmax.npy"), hi)
        args += ["-m", str(tmp_path / "dmin.npy"), "-M", str(tmp_path / "dmax.npy")]
        hi = np.where(hi < lo + 1, np.ceil(lo + 1), hi).astype(np.float32)   # the repair of mgm.cc:349-352
    else:
        lo, hi = np.full((ny, nx), -23, np.float32), np.full((ny, nx), 4, np.float32)
    names = [str(tmp_path / ("ref_%s.npy" % k)) for k in ("disp", "cost")]
    env = dict(os.environ, TSGM=str(K), TSGM_ITER=str(iters), TESTLRRL="0", USE_TRUNCATED_LINEAR_POTENTIALS=str(felz))
    r = subprocess.run([REF_MGM] + args + [str(tmp_path / "u.pnm"), str(tmp_path / "v.pnm")] + names, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    disp, cost = [np.squeeze(np.load(n)) for n in names]
    uu, vv = u.astype(np.float32), v.astype(np.float32)
    emin, emax = int(lo.min()) - 3 * (iters - 1), int(hi.max()) + 3 * (iters - 1)
    P1, P2 = (2.0, 20000.0) if felz else (8.0, 32.0)
    w = O.orc_weights(uu, 1.0, 5.0)
    cc = O.orc_costvolume_ranges(uu, vv, lo, hi, emin, emax, "none", "ad", 40.0, 3)
    slo, shi = lo.copy(), hi.copy()
    for _ in range(iters):
        rr = O.orc_mgm_ranges(cc, lo, hi, w, emin, slo, shi, P1, P2, 8, K, felz, 1)
        o, oc = O.orc_refine_ranges(rr["S"], slo, shi, emin, rr["out"], rr["outcost"], "vfit")
        slo, shi, (gmin, gmax) = O.orc_update_range(o, slo, shi)
        slo[~np.isfinite(slo)] = gmin
        shi[~np.isfinite(shi)] = gmax
    ok = np.isfinite(cost)   # the reference leaves the label uninitialised where no label of the range is finite
    assert same(cost, oc) and same(disp[ok], o[ok]) and ok.mean() > 0.25


# ------------------------------------------------------------------------------------------ golden flows
def test_flow_interface_matches_api():
    """OracleImpl (tests/flow_checks.py) takes exactly the keywords of mgm_b200.api.Context for the two flow calls, so
    that the same checks drive both."""
    import inspect
    from mgm_b200.api import Context
    from tests.flow_checks import OracleImpl
    for name in ("stereo_lr", "stereo_ranges"):
        a = inspect.signature(getattr(Context, name)).parameters
        b = inspect.signature(getattr(OracleImpl, name)).parameters
        assert list(a) == list(b), (name, list(a), list(b))
        for k in a:
            if a[k].default is not inspect.Parameter.empty:
                assert repr(a[k].default) == repr(b[k].default), (name, k)


@pytest.mark.parametrize("path", golden_files("flow_lr"))
def test_golden_flow_lr_oracle(path):
    from tests.flow_checks import OracleImpl, check_flow_lr
    check_flow_lr(load_golden(path), OracleImpl())


@pytest.mark.parametrize("path", golden_files("flow_ranges"))
def test_golden_flow_ranges_oracle(path):
    from tests.flow_checks import OracleImpl, check_flow_ranges
    check_flow_ranges(load_golden(path), OracleImpl())


# ------------------------------------------------------------------------------------------ sweeps 8-15 (-O 16)
def test_knight_sweeps_definition():
    """Sweeps 8-15 are defined by this build (the reference's table ends at 8, mgm_core.cc:463-473).  The definition
    (mgm_b200/csrc/common.cuh, oracle/mgm_oracle.c orc_pass_order): same scan and same four neighbours as sweep b = p-8,
    order by coordinate parity such that the first-neighbour chain advances by a knight move every two pixels and the
    second-neighbour chain by the perpendicular knight move."""
    import ctypes
    lib = O.orc_lib()
    def nb(p, xs, ys, k):
        dx, dy = ctypes.c_int(), ctypes.c_int()
        lib.orc_pass_neighbour(p, xs, ys, k, ctypes.byref(dx), ctypes.byref(dy))
        return dx.value, dy.value
    nx, ny = 9, 7
    expect_k0 = {8: (-2, -1), 9: (2, 1), 10: (-1, 2), 11: (1, -2), 12: (-1, -2), 13: (2, -1), 14: (1, 2), 15: (-2, 1)}
    for p in range(8, 16):
        base = [nb(p - 8, 3, 3, k) for k in range(4)]
        for xs in range(2, 6):
            for ys in range(2, 6):
                cur = [nb(p, xs, ys, k) for k in range(4)]
                assert sorted(cur) == sorted(base)          # the same four neighbours, reordered
                # two steps along neighbour k: scan coordinates of the neighbour (from the base table in scan space)
                pre = O.orc_scan_preds(p - 8, nx, ny)       # scan-space offsets of the base order
                for k in (0, 1):
                    d1 = cur[k]
                    s1 = pre[base.index(d1)]                # scan offset of that neighbour
                    d2 = nb(p, xs + s1[0], ys + s1[1], k)
                    step2 = (d1[0] + d2[0], d1[1] + d2[1])
                    assert abs(step2[0]) + abs(step2[1]) == 3 and 0 not in step2, (p, xs, ys, k, step2)
                    if k == 0:
                        assert step2 == expect_k0[p], (p, step2)
                    else:
                        assert step2[0] * expect_k0[p][0] + step2[1] * expect_k0[p][1] == 0   # perpendicular
    # with all four neighbours the sweeps 8+b and b differ by the order of one float sum only; with fewer they differ
    cc = synth_volume(23, 17, 9, seed=4, inf_border=False)
    r = O.orc_mgm(cc, None, 0, 8, 32, 16, 4, 0, 1, want_passes=True)
    for b in range(8):
        assert np.allclose(r["passes"][8 + b], r["passes"][b], rtol=1e-5)
    r = O.orc_mgm(cc, None, 0, 8, 32, 16, 2, 0, 1, want_passes=True)
    assert all(not np.allclose(r["passes"][8 + b], r["passes"][b], rtol=1e-3) for b in range(8))


def test_census_prefilter_with_another_distance():
    """"-p census -t ad|sd|ncc|btad|btsd": the reference picks the cost function before it forces census/census
    (mgm_costvolume.h:355 vs :358-362), so the distance runs on the census bit strings held in float channels
    (denormals, and with 7x7 windows NaN / INF bit patterns).  The port reproduces it bit for bit."""
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    for nch in (1, 3):
        u, v = synth_pair(45, 29, 14, seed=4, nch=nch)
        for win in (3, 5, 7):
            for dist in ("ad", "sd", "ncc", "btad", "btsd"):
                for trunc in (np.inf, 1e-38):
                    a = O.orc_costvolume(u, v, -13, 3, "census", dist, trunc, win)
                    b = O.ref_costvolume(u, v, -13, 3, "census", dist, trunc, win)
                    assert np.array_equal(a, b, equal_nan=True), (nch, win, dist, trunc)


def test_nonfinite_volumes_port_vs_reference():
    """The port against the reference on volumes with all-INF vectors, NaN and -INF costs, every update variant: this pins
    the checker used by the GPU test of the compare-select kernel (tests/test_gpu_parity.py::test_mgm_nonfinite_volumes).
    Where no label is finite the reference's disparity is uninitialised (mgm_core.cc:594): compared where the cost is."""
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5)
    for (nx, ny, L) in [(23, 17, 9), (31, 22, 12)]:
        for kind in ("allinf", "nan", "neginf", "mix"):
            cc = synth_volume(nx, ny, L, seed=3, real=True)
            if kind in ("allinf", "mix"):
                for _ in range(6):
                    cc[rng.integers(ny), rng.integers(nx), :] = np.inf
            if kind in ("nan", "mix"):
                for _ in range(8):
                    cc[rng.integers(ny), rng.integers(nx), rng.integers(L)] = np.nan
            if kind in ("neginf", "mix"):
                for _ in range(4):
                    cc[rng.integers(ny), rng.integers(nx), rng.integers(L)] = -np.inf
            for K, felz, wt in itertools.product((1, 2, 3, 4), (0, 1), (0, 1)):
                w = synth_weights(nx, ny, seed=K) if wt else None
                P1, P2 = (8, 32) if not felz else (2, 20000)
                a = O.orc_mgm(cc, w, -(L - 1), P1, P2, 8, K, felz, 1)
                b = O.ref_mgm(cc, w, -(L - 1), P1, P2, 8, K, felz, 1)
                fin = np.isfinite(b["outcost"])
                assert np.array_equal(a["S"], b["S"], equal_nan=True), (kind, K, felz, wt)
                assert np.array_equal(a["out"][fin], b["out"][fin]) and np.array_equal(a["outcost"], b["outcost"], equal_nan=True)
                assert np.isnan(a["out"][~fin]).all()
    # non-finite / negative penalties and weights
    cc = synth_volume(31, 22, 12, seed=8, real=True)
    w = synth_weights(31, 22, seed=2)
    w[3, 5, 7] = -2.0
    w[1, 9, 9] = np.inf
    for K, felz, P1, P2 in [(2, 0, 8, np.inf), (3, 0, np.inf, 32), (2, 1, -1.0, 20), (4, 1, 2, 20000)]:
        for ww in (None, w):
            a = O.orc_mgm(cc, ww, -11, P1, P2, 8, K, felz, 1)
            b = O.ref_mgm(cc, ww, -11, P1, P2, 8, K, felz, 1)
            fin = np.isfinite(b["outcost"])
            assert np.array_equal(a["S"], b["S"], equal_nan=True) and np.array_equal(a["out"][fin], b["out"][fin])
