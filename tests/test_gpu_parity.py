"""GPU suite (-m gpu): the CUDA path, called through the C ABI (mgm_b200.Context -> libmgmb200.so), against
 * the oracle port on seeded inputs at sizes it finishes in seconds          (bit-exact, see below)
 * the golden vectors generated from the unmodified reference               (bit-exact)
 * size-independent properties at BASELINE.json's full sizes.
Tolerances: WTA disparity indices, aggregated costs S, output costs and sub-pixel offsets are all compared
BIT-EXACT here (the kernels perform the reference's IEEE operations in the reference's order; this is
stricter than north_star's 1e-4 relative).  NaN patterns must coincide (equal_nan)."""
import itertools

import numpy as np
import pytest

import oracle as O
from tests.conftest import synth_pair, synth_volume, synth_weights
from tests.golden_util import golden_files, load_golden

pytestmark = pytest.mark.gpu


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


def mism(a, b):
    return float(np.mean(~((a == b) | (np.isnan(a) & np.isnan(b)))))


# ------------------------------------------------------------------------------------------ stages
@pytest.mark.parametrize("nch", [1, 3])
def test_weights(ctx, nch):
    u, _ = synth_pair(53, 31, 12, seed=3, nch=nch)
    for aP, aT in [(4.0, 12.0), (0.5, 40.0), (1.0, 5.0)]:
        assert same(ctx.compute_mgm_weights(u, aP, aT), O.orc_weights(u, aP, aT))


@pytest.mark.parametrize("dist,win", [("ad", 3), ("sd", 3), ("census", 3), ("census", 5), ("census", 7), ("ncc", 3),
                                      ("ncc", 5), ("ncc", 7), ("btad", 3), ("btsd", 3)])
@pytest.mark.parametrize("nch", [1, 3])
def test_costvolume(ctx, dist, win, nch):
    u, v = synth_pair(45, 29, 14, seed=win + nch, nch=nch)
    u = u + np.float32(0.37)
    for trunc in [np.inf, 17.5]:
        a = ctx.allocate_and_fill_sgm_costvolume(u, v, -13, 3, "none", dist, trunc, win)
        assert same(a, O.orc_costvolume(u, v, -13, 3, "none", dist, trunc, win)), (dist, win, nch, trunc)
    a = ctx.allocate_and_fill_sgm_costvolume(u, v[:, :, :33], -13, 3, "none", dist, np.inf, win)   # ragged sizes
    assert same(a, O.orc_costvolume(u, v[:, :, :33], -13, 3, "none", dist, np.inf, win))


def test_costvolume_census_prefilter_quirk(ctx):
    """-p census with a non-census distance (mgm_costvolume.h:355 vs :358-362): the distance runs on the census bit
    strings read as floats (denormals; NaN / INF patterns with 7x7 windows) -- like the reference, bit for bit."""
    for nch in (1, 3):
        u, v = synth_pair(45, 29, 14, seed=4, nch=nch)
        for win in (3, 5, 7):
            for dist in ("ad", "sd", "ncc", "btad", "btsd"):
                for trunc in (np.inf, 1e-38):
                    a = ctx.allocate_and_fill_sgm_costvolume(u, v, -13, 3, "census", dist, trunc, win)
                    b = O.orc_costvolume(u, v, -13, 3, "census", dist, trunc, win)
                    assert same(a, b), (nch, win, dist, trunc, mism(a, b))


def test_costvolume_edge_cases(ctx):
    u, v = synth_pair(40, 9, 60, seed=1)
    # range entirely outside the right image for most pixels -> the all-invalid rule (mgm_costvolume.h:414-421)
    a = ctx.allocate_and_fill_sgm_costvolume(u, v, 30, 75, "none", "ad", np.inf, 3)
    b = O.orc_costvolume(u, v, 30, 75, "none", "ad", np.inf, 3)
    assert same(a, b) and (b[:, -5:, :] == 0).all()
    # sobelx prefilter + truncation, unknown names fall back to entry 0
    assert same(ctx.allocate_and_fill_sgm_costvolume(u, v, -20, 3, "sobelx", "sd", 90.0, 3),
                O.orc_costvolume(u, v, -20, 3, "sobelx", "sd", 90.0, 3))
    assert same(ctx.allocate_and_fill_sgm_costvolume(u, v, -20, 3, "sobel_x", "l1", np.inf, 3),
                O.orc_costvolume(u, v, -20, 3, "none", "ad", np.inf, 3))


def test_costvolume_ncc_thread_per_pixel(ctx):
    """Single-channel NCC (mgm_costvolume_ncc1_kernel: one thread per pixel, register ring of window columns, staged
    rows, transposed stores): rows wider than a block of 128 pixels, label counts that are not a multiple of 32, NaN
    samples, matches outside the right image, truncation, the all-invalid rule, per-pixel ranges."""
    rng = np.random.default_rng(5)
    for (nx, ny, L, win) in [(300, 9, 70, 5), (131, 12, 33, 3), (260, 11, 40, 7)]:
        u, v = synth_pair(nx, ny, L, seed=win, nch=1)
        u = u + np.float32(0.21)
        for _ in range(6):
            u[0, rng.integers(ny), rng.integers(nx)] = np.nan
            v[0, rng.integers(ny), rng.integers(nx)] = np.nan
        for dmin, trunc in [(-(L - 10), np.inf), (-5, 30.0), (nx - 20, np.inf)]:
            a = ctx.allocate_and_fill_sgm_costvolume(u, v, dmin, dmin + L - 1, "none", "ncc", trunc, win)
            b = O.orc_costvolume(u, v, dmin, dmin + L - 1, "none", "ncc", trunc, win)
            assert same(a, b), (nx, ny, L, win, dmin, trunc, mism(a, b))
        lo, hi = _ragged(nx, ny, -(L - 1), 0, 2)
        assert same(ctx.costvolume_ranges(u, v, lo, hi, -(L - 1), 0, "none", "ncc", np.inf, win),
                    O.orc_costvolume_ranges(u, v, lo, hi, -(L - 1), 0, "none", "ncc", np.inf, win)), (nx, win)


@pytest.mark.parametrize("nch", [1, 3])
def test_costvolume_gblur_prefilter(ctx, nch):
    """-p gblur (truncated Gaussian, sigma 1, Neumann borders) before AD / SD / BT / NCC; ignored by census"""
    u, v = synth_pair(57, 33, 14, seed=20 + nch, nch=nch)
    u = u + np.float32(0.41)
    for dist, trunc in [("ad", np.inf), ("sd", 400.0), ("btsd", np.inf), ("ncc", np.inf), ("census", np.inf)]:
        a = ctx.allocate_and_fill_sgm_costvolume(u, v, -13, 3, "gblur", dist, trunc, 3)
        assert same(a, O.orc_costvolume(u, v, -13, 3, "gblur", dist, trunc, 3)), (dist, nch)
    a = ctx.allocate_and_fill_sgm_costvolume(u[:, :3, :2], v[:, :3, :2], -1, 1, "gblur", "ad", np.inf, 3)   # image smaller than the kernel
    assert same(a, O.orc_costvolume(u[:, :3, :2], v[:, :3, :2], -1, 1, "gblur", "ad", np.inf, 3))


MGM_SHAPES = [(23, 17, 9), (67, 41, 19), (131, 37, 40)]


@pytest.mark.parametrize("K", [1, 2, 3, 4])
@pytest.mark.parametrize("felz", [0, 1])
@pytest.mark.parametrize("weighted", [0, 1])
def test_mgm_matches_oracle(ctx, K, felz, weighted):
    """all message-update variants (mgm_core.cc:66-281) x sweep counts x band layouts, S/out/outcost bit-exact"""
    for (nx, ny, L), real, rows in itertools.product(MGM_SHAPES, [0, 1], [0, 5]):
        cc = synth_volume(nx, ny, L, seed=nx + K, real=bool(real))
        w = synth_weights(nx, ny, seed=nx) if weighted else None
        ctx.set_rows_per_band(rows)   # 5 rows per band forces many chained bands on these small images
        for NDIR in ([8] if rows == 0 else [1, 4, 8]):
            P1, P2 = ((8, 32) if K % 2 else (3.5, 21.25)) if not felz else ((2, 20000) if K != 4 else (1.5, 11))
            r = ctx.mgm(cc, w, -(L - 1), P1, P2, NDIR, K, felz, 1)
            o = O.orc_mgm(cc, w, -(L - 1), P1, P2, NDIR, K, felz, 1)
            tag = (nx, ny, L, real, rows, NDIR, mism(r["S"], o["S"]), mism(r["out"], o["out"]))
            assert same(r["out"], o["out"]), tag
            assert same(r["S"], o["S"]), tag
            assert same(r["outcost"], o["outcost"]), tag
    ctx.set_rows_per_band(0)


@pytest.mark.parametrize("knob", ["MGMB200_GROUPS=2", "MGMB200_GROUPS=3", "MGMB200_NO_SHEAR=1", "MGMB200_NO_CREG=1",
                                  "MGMB200_STATIC_ORDER=1", "MGMB200_NO_FUSED_SGM=1", "MGMB200_LANES4=1",
                                  "MGMB200_NO_FUSED_FINISH=1", "MGMB200_FUSED_FINISH=1", "MGMB200_FIN_TILE=7x3", "MGMB200_FIN_TILE=4096x4096",
                                  "MGMB200_CC_PF=3", "MGMB200_REG_CHAINS=1", "MGMB200_NO_LEAN_SGM=1", "MGMB200_NO_LEAN_TRUNC=1"])
def test_mgm_alternative_kernel_layouts(ctx, knob):
    """The aggregation kernel's alternative layouts (row groups on their own named barriers, row-per-worker diagonal
    sweeps, cp.async cost ring, static band order, finish stage as a separate launch or fused with other tile sizes,
    L2 prefetch of the costs) are selected by context options (mgmb200_set_option; the MGMB200_* environment is only
    read when a context is created): each must give the same bits as the default layout, i.e. as the oracle."""
    name, val = knob.split("=")
    ctx.set_option("fused_finish", 1)   # these small label counts would not take the finish tiles by default
    ctx.set_option(name[len("MGMB200_"):].lower(), val)
    try:
        for (nx, ny, L), (K, felz, P1, P2) in itertools.product([(131, 37, 40), (90, 150, 24)],
                                                                [(3, 1, 2, 20000), (2, 1, 2, 20000), (2, 0, 8, 32), (3, 0, 8, 32)]):
            cc = synth_volume(nx, ny, L, seed=nx + K, real=True)
            for rows in (0, 16):
                ctx.set_rows_per_band(rows)
                r = ctx.mgm(cc, None, -(L - 1), P1, P2, 8, K, felz, 1)
                o = O.orc_mgm(cc, None, -(L - 1), P1, P2, 8, K, felz, 1)
                assert same(r["S"], o["S"]) and same(r["out"], o["out"]), (knob, nx, ny, L, K, felz, rows, mism(r["S"], o["S"]))
    finally:
        ctx.set_option("reset")


@pytest.mark.parametrize("lanes", [0, 4])
def test_mgm_lean_sgm_kernels(ctx, lanes):
    """The lean unweighted-SGM kernels (aggregate_sgm.cu: compile-time label layout, split warp roles, double-buffered
    costs) against the oracle for every chunk count per lane they are built for (2, 4, 6, 8 with 8 lanes per worker;
    4, 8 with 4 lanes), TSGM 1-4, chained bands (16 rows) and default bands, with and without the L2 cost prefetch;
    the generic kernel runs the same cases under MGMB200_NO_LEAN_SGM=1 and must give the same bits."""
    labels = [40, 100, 180, 250] if lanes == 0 else [50, 120]
    try:
        if lanes:
            ctx.set_option("lanes4", 1)
        for L, K in itertools.product(labels, (1, 2, 3, 4)):
            nx, ny = (75, 44) if L > 150 else (101, 58)
            cc = synth_volume(nx, ny, L, seed=L + K, real=True)
            o = O.orc_mgm(cc, None, -(L - 1), 8, 32, 8, K, 0, 1)
            for rows, pf, lean in ((16, 3, 1), (0, 0, 1), (16, 3, 0)):
                ctx.set_rows_per_band(rows)
                ctx.set_option("cc_pf", pf)
                ctx.set_option("no_lean_sgm", 0 if lean else 1)
                r = ctx.mgm(cc, None, -(L - 1), 8, 32, 8, K, 0, 1)
                tag = (lanes, L, K, rows, pf, lean, mism(r["S"], o["S"]), mism(r["out"], o["out"]))
                assert same(r["S"], o["S"]) and same(r["out"], o["out"]) and same(r["outcost"], o["outcost"]), tag
    finally:
        ctx.set_rows_per_band(0)
        ctx.set_option("reset")


def test_mgm_lean_weighted_sgm_kernels(ctx):
    """The lean SGM kernels with per-edge weights (aggregate_sgmw.cu: raw messages in the ring, consumer-side transform with
    the weights of the pixel, every sweep row-per-worker) against the oracle: chunk counts 2, 4, 6, 8 per lane, TSGM 1-4,
    chained bands (16 rows) and default bands; the generic kernel runs the same cases under MGMB200_NO_LEAN_SGM=1."""
    try:
        for L, K in itertools.product([40, 100, 180, 250], (1, 2, 3, 4)):
            nx, ny = (75, 44) if L > 150 else (101, 58)
            cc = synth_volume(nx, ny, L, seed=L + K, real=True)
            w = synth_weights(nx, ny, seed=K)
            o = O.orc_mgm(cc, w, -(L - 1), 8, 32, 8, K, 0, 1)
            for rows, lean in ((16, 1), (0, 1), (16, 0)):
                ctx.set_rows_per_band(rows)
                ctx.set_option("no_lean_sgm", 0 if lean else 1)
                r = ctx.mgm(cc, w, -(L - 1), 8, 32, 8, K, 0, 1)
                tag = (L, K, rows, lean, mism(r["S"], o["S"]), mism(r["out"], o["out"]))
                assert same(r["S"], o["S"]) and same(r["out"], o["out"]) and same(r["outcost"], o["outcost"]), tag
        # thin images and tiny bands
        for nx, ny in [(2, 9), (9, 2), (1, 1), (6, 1), (3, 3), (57, 3), (3, 57)]:
            cc = synth_volume(nx, ny, 40, seed=nx + ny, real=True)
            w = synth_weights(nx, ny, seed=nx)
            for K, rows in itertools.product((2, 4), (0, 1, 2)):
                ctx.set_rows_per_band(rows)
                r = ctx.mgm(cc, w, -39, 8, 32, 8, K, 0, 1)
                o = O.orc_mgm(cc, w, -39, 8, 32, 8, K, 0, 1)
                assert same(r["S"], o["S"]) and same(r["out"], o["out"]), (nx, ny, K, rows, mism(r["S"], o["S"]))
    finally:
        ctx.set_rows_per_band(0)
        ctx.set_option("reset")


def test_mgm_lean_trunc_kernels(ctx):
    """The lean unweighted truncated-linear kernels (aggregate_trunc.cu: compile-time label layout, band hand-off off the
    step barriers) against the oracle for every chunk count per lane they are built for (2, 4, 6, 8), TSGM 1-4, chained
    bands (16 rows) and default bands; the generic kernel runs the same cases under MGMB200_NO_LEAN_TRUNC=1."""
    try:
        for L, K in itertools.product([40, 100, 180, 250], (1, 2, 3, 4)):
            nx, ny = (75, 44) if L > 150 else (101, 58)
            cc = synth_volume(nx, ny, L, seed=L + K, real=True)
            P1, P2 = (2, 20000) if K != 4 else (1.5, 11)
            o = O.orc_mgm(cc, None, -(L - 1), P1, P2, 8, K, 1, 1)
            for rows, lean in ((16, 1), (0, 1), (16, 0)):
                ctx.set_rows_per_band(rows)
                ctx.set_option("no_lean_trunc", 0 if lean else 1)
                r = ctx.mgm(cc, None, -(L - 1), P1, P2, 8, K, 1, 1)
                tag = (L, K, rows, lean, mism(r["S"], o["S"]), mism(r["out"], o["out"]))
                assert same(r["S"], o["S"]) and same(r["out"], o["out"]) and same(r["outcost"], o["outcost"]), tag
    finally:
        ctx.set_rows_per_band(0)
        ctx.set_option("reset")


def test_mgm_lean_kernels_degenerate_sizes(ctx):
    """Images with one or two rows / columns, single pixels and thin strips through the lean kernels (label counts they
    take: 40 labels = 2 chunks per lane), both potentials, TSGM 1-4, 16 sweeps for SGM, chained bands of 2 rows."""
    try:
        for nx, ny in [(2, 9), (9, 2), (1, 1), (1, 6), (6, 1), (2, 2), (3, 3), (4, 3), (3, 2), (57, 3), (3, 57), (70, 1), (1, 70)]:
            cc = synth_volume(nx, ny, 40, seed=nx + 3 * ny, real=True)
            for felz, K, NDIR, rows in itertools.product((0, 1), (1, 2, 3, 4), (8, 16), (0, 2)):
                if NDIR == 16 and (felz or K not in (2, 4)):
                    continue
                ctx.set_rows_per_band(rows)
                P1, P2 = (8, 32) if not felz else (2, 20000)
                r = ctx.mgm(cc, None, -39, P1, P2, NDIR, K, felz, 1)
                o = O.orc_mgm(cc, None, -39, P1, P2, NDIR, K, felz, 1)
                assert same(r["S"], o["S"]) and same(r["out"], o["out"]), (nx, ny, felz, K, NDIR, rows, mism(r["S"], o["S"]))
        # bands of 1, 3 and 5 rows (fewer rows than a warp holds: the chains still need two compute warps)
        cc = synth_volume(40, 30, 40, seed=1, real=True)
        for felz, K, rows in itertools.product((0, 1), (1, 2, 3, 4), (1, 3, 5)):
            ctx.set_rows_per_band(rows)
            P1, P2 = (8, 32) if not felz else (2, 20000)
            r = ctx.mgm(cc, None, -39, P1, P2, 8, K, felz, 1)
            o = O.orc_mgm(cc, None, -39, P1, P2, 8, K, felz, 1)
            assert same(r["S"], o["S"]) and same(r["out"], o["out"]), (felz, K, rows, mism(r["S"], o["S"]))
    finally:
        ctx.set_rows_per_band(0)


def test_mgm_overcount_flag_and_small_images(ctx):
    cc = synth_volume(31, 22, 12, seed=8, real=True)
    for fix in [0, 1]:
        r = ctx.mgm(cc, None, -11, 8, 32, 8, 2, 0, fix)
        o = O.orc_mgm(cc, None, -11, 8, 32, 8, 2, 0, fix)
        assert same(r["S"], o["S"]) and same(r["out"], o["out"])
    for nx, ny in [(2, 9), (9, 2), (1, 1), (1, 6), (6, 1), (2, 2), (3, 3), (4, 3), (3, 2)]:   # degenerate sizes: few or no interior pixels
        cc = synth_volume(nx, ny, 7, seed=nx, inf_border=False)
        r = ctx.mgm(cc, None, 0, 8, 32, 8, 4, 0, 1)
        o = O.orc_mgm(cc, None, 0, 8, 32, 8, 4, 0, 1)
        assert same(r["S"], o["S"]) and same(r["out"], o["out"]), (nx, ny)


def test_mgm_label_counts(ctx):
    """label counts around the padding granularity, two labels (minimum), a few hundred"""
    for L in [2, 3, 31, 32, 33, 64, 65, 300]:
        cc = synth_volume(37, 11, L, seed=L, real=True, inf_border=(L < 37))
        for K, felz in [(2, 0), (3, 1)]:
            r = ctx.mgm(cc, None, -5, 8, 32, 4, K, felz, 1)
            o = O.orc_mgm(cc, None, -5, 8, 32, 4, K, felz, 1)
            assert same(r["S"], o["S"]) and same(r["out"], o["out"]), (L, K, felz)


def test_mgm_labelmajor_protocol(ctx):
    """matlab/mgm_o.cc input.bin layout: costs[i + o*ncol*nrow], labels 0..nlab-1"""
    cc = synth_volume(41, 23, 16, seed=3, inf_border=False)
    w = synth_weights(41, 23, seed=3)
    lab, _ = ctx.mgm_labelmajor(np.ascontiguousarray(np.transpose(cc, (2, 0, 1))), w, 8, 32, 8, 2, 0)
    o = O.orc_mgm(cc, w, 0, 8, 32, 8, 2, 0, 1)
    assert same(lab, o["out"])


def _nonfinite_volume(nx, ny, L, kind, seed):
    rng = np.random.default_rng(seed)
    cc = synth_volume(nx, ny, L, seed=seed, real=True)
    if kind in ("allinf", "mix"):
        for _ in range(6):
            cc[rng.integers(ny), rng.integers(nx), :] = np.inf      # a pixel without any finite cost
    if kind in ("nan", "mix"):
        for _ in range(8):
            cc[rng.integers(ny), rng.integers(nx), rng.integers(L)] = np.nan
    if kind in ("neginf", "mix"):
        for _ in range(4):
            cc[rng.integers(ny), rng.integers(nx), rng.integers(L)] = -np.inf
    return cc


@pytest.mark.parametrize("kind", ["allinf", "nan", "neginf", "mix"])
def test_mgm_nonfinite_volumes(ctx, kind):
    """User-supplied volumes with all-INF vectors, NaN or -INF costs (the mgm(CC, ...) / mgm_o boundary): the reference
    propagates them by its compare-select minima and '<' scans (mgm_core.cc:47-60, dvec.cc:81-88).  The
    host-pointer entry points detect such inputs and run the compare-select kernel (aggregate_generic.cu): same bits as
    the oracle port, which tests/test_oracle.py pins against the reference on the same kind of volumes.  Pixels without
    a finite label are NaN here (uninitialised in the reference, mgm_core.cc:594)."""
    for (nx, ny, L) in [(23, 17, 9), (45, 31, 40)]:
        cc = _nonfinite_volume(nx, ny, L, kind, 3)
        for K, felz, wt, NDIR in itertools.product((1, 2, 3, 4), (0, 1), (0, 1), (8, 16)):
            if NDIR == 16 and (K, wt) not in ((2, 0), (3, 1)):
                continue
            w = synth_weights(nx, ny, seed=K) if wt else None
            P1, P2 = (8, 32) if not felz else (2, 20000)
            r = ctx.mgm(cc, w, -(L - 1), P1, P2, NDIR, K, felz, 1)
            o = O.orc_mgm(cc, w, -(L - 1), P1, P2, NDIR, K, felz, 1)
            assert same(r["S"], o["S"]) and same(r["out"], o["out"]) and same(r["outcost"], o["outcost"]), \
                (kind, nx, K, felz, wt, NDIR, mism(r["S"], o["S"]))
    # non-finite penalties and weights take the same path
    cc = synth_volume(31, 22, 12, seed=8, real=True)
    w = synth_weights(31, 22, seed=2)
    w[3, 5, 7] = -2.0
    w[1, 9, 9] = np.inf
    for K, felz, P1, P2 in [(2, 0, 8, np.inf), (3, 0, np.inf, 32), (2, 1, -1.0, 20), (4, 1, 2, 20000)]:
        for ww in (None, w):
            r = ctx.mgm(cc, ww, -11, P1, P2, 8, K, felz, 1)
            o = O.orc_mgm(cc, ww, -11, P1, P2, 8, K, felz, 1)
            assert same(r["S"], o["S"]) and same(r["out"], o["out"]), (K, felz, P1, P2, ww is not None)
    # the matlab/mgm_o.cc protocol entry
    cc = _nonfinite_volume(41, 23, 16, "mix", 9)
    lab, _ = ctx.mgm_labelmajor(np.ascontiguousarray(np.transpose(cc, (2, 0, 1))), None, 8, 32, 8, 2, 0)
    assert same(lab, O.orc_mgm(cc, None, 0, 8, 32, 8, 2, 0, 1)["out"])


def test_unsupported_inputs_fail_loudly(ctx):
    import mgm_b200
    cc = synth_volume(9, 9, 5, seed=1, inf_border=False)
    with pytest.raises(mgm_b200.MgmError):
        ctx.mgm(cc, None, 0, 8, 32, 17, 2)             # 16 sweeps are defined (8 of the reference + 8 of this build), not more
    with pytest.raises(mgm_b200.MgmError):
        ctx.mgm(cc, None, 0, 8, 32, 4, 5)
    bad = cc.copy(); bad[2, 3, 1] = np.nan              # per-pixel ranges stay on the fast kernels: non-finite costs are refused
    with pytest.raises(mgm_b200.MgmError):
        ctx.mgm_ranges(bad, 0, 4, None, 0, 0, 4, 8, 32, 4, 2)


@pytest.mark.parametrize("method", ["none", "vfit", "parabola", "cubic", "parabolaOCV", "bogus"])
def test_refinement(ctx, method):
    cc = synth_volume(67, 41, 19, seed=5, real=True)
    o = O.orc_mgm(cc, None, -18, 8, 32, 8, 2, 0, 1)
    a = ctx.subpixel_refinement_sgm(o["S"], -18, o["out"], o["outcost"], method)
    b = O.orc_refine(o["S"], -18, o["out"], o["outcost"], method)
    assert same(a[0], b[0]) and same(a[1], b[1])


def test_sweep_sharding_api_single_gpu(ctx):
    """The multi-GPU protocol on one device: aggregate the sweeps in two separate calls (as two ranks would),
    finish two row slabs with explicit sweep pointers -> identical to the one-call path and to the oracle."""
    import torch
    nx, ny, L = 83, 47, 24
    cc = synth_volume(nx, ny, L, seed=21, real=True)
    VS = ctx.padded_labels(L)
    dense = torch.from_numpy(cc).cuda()
    dcc = torch.empty((ny, nx, VS), device="cuda")
    dout = torch.full((ny, nx), -777.0, device="cuda")
    dcost = torch.empty((ny, nx), device="cuda")
    ctx.pad_volume_dev(dense.data_ptr(), dcc.data_ptr(), nx, ny, L)
    for mask in (0x55, 0xAA):
        ctx.aggregate_sweeps_dev(dcc.data_ptr(), 0, 0, nx, ny, -(L - 1), 0, 2.0, 20000.0, 8, 3, 1, mask)
    ptrs = [ctx.sweep_volume(p)[0] for p in range(8)]
    for r0, r1 in ((0, 20), (20, ny)):
        ctx.finish_rows_dev(ptrs, dcc.data_ptr(), nx, ny, -(L - 1), 0, 8, 1, "vfit", r0, r1, dout.data_ptr(), dcost.data_ptr())
    ctx.synchronize()
    o = O.orc_mgm(cc, None, -(L - 1), 2.0, 20000.0, 8, 3, 1, 1)
    ro, rc = O.orc_refine(o["S"], -(L - 1), o["out"], o["outcost"], "vfit")
    assert same(dout.cpu().numpy(), ro) and same(dcost.cpu().numpy(), rc)


def test_sweep_slab_stores_single_gpu(ctx):
    """The ordered multi-GPU exchange on one device: the aggregation kernel stores the messages of the lower row slab
    into a second set of volumes (what a peer mapping would be), every slab is finished from "its" volumes in sweep
    order -> identical to the one-call path, for row-major and column-major sweeps, sheared and knight sweeps."""
    import torch
    for (nx, ny, L, NDIR, K, felz, P1, P2, nslabs) in [(83, 47, 24, 8, 3, 1, 2.0, 20000.0, 2), (61, 90, 40, 16, 2, 0, 8.0, 32.0, 3),
                                                       (50, 33, 12, 8, 4, 0, 8.0, 32.0, 2)]:
        cc = synth_volume(nx, ny, L, seed=nx, real=True)
        VS = ctx.padded_labels(L)
        dense = torch.from_numpy(cc).cuda()
        dcc = torch.empty((ny, nx, VS), device="cuda")
        dout = torch.full((ny, nx), -777.0, device="cuda")
        dcost = torch.empty((ny, nx), device="cuda")
        ctx.pad_volume_dev(dense.data_ptr(), dcc.data_ptr(), nx, ny, L)
        ctx.sweeps_alloc(nx, ny, -(L - 1), 0, NDIR)
        rps = max(2, -(-ny // nslabs))
        own = [ctx.sweep_volume(p)[0] for p in range(NDIR)]
        other = [[torch.full((ny, nx, VS), float("nan"), device="cuda") for p in range(NDIR)] for r in range(nslabs - 1)]
        table = [[own[p]] + [other[r][p].data_ptr() for r in range(nslabs - 1)] for p in range(NDIR)]
        torch.cuda.synchronize()   # torch's stream and the context's are not ordered with each other
        for rows in (0, 7):
            ctx.set_rows_per_band(rows)
            for mask in ((0x5555, 0xAAAA) if rows else (0xFFFF,)):
                ctx.aggregate_sweeps_slabs_dev(dcc.data_ptr(), 0, 0, nx, ny, -(L - 1), 0, P1, P2, NDIR, K, felz, mask, nslabs,
                                               rps, table)
            for r in range(nslabs):
                r0, r1 = min(ny, r * rps), min(ny, (r + 1) * rps)
                ctx.finish_rows_dev([table[p][r] for p in range(NDIR)], dcc.data_ptr(), nx, ny, -(L - 1), 0, NDIR, 1, "vfit",
                                    r0, r1, dout.data_ptr(), dcost.data_ptr())
            ctx.synchronize()
            o = O.orc_mgm(cc, None, -(L - 1), P1, P2, NDIR, K, felz, 1)
            ro, rc = O.orc_refine(o["S"], -(L - 1), o["out"], o["outcost"], "vfit")
            assert same(dout.cpu().numpy(), ro) and same(dcost.cpu().numpy(), rc), (nx, ny, L, NDIR, K, felz, rows)
        ctx.set_rows_per_band(0)
        # the all-reduce exchange on one device: partial sums of two "ranks", added, finished from the sum
        part = [torch.empty((ny, nx, VS), device="cuda") for _ in range(2)]
        ctx.aggregate_sweeps_dev(dcc.data_ptr(), 0, 0, nx, ny, -(L - 1), 0, P1, P2, NDIR, K, felz, (1 << NDIR) - 1)
        for i, mask in enumerate((0x5555, 0xAAAA)):
            ctx.sum_sweeps_dev(nx, ny, -(L - 1), 0, mask & ((1 << NDIR) - 1), part[i].data_ptr())
        ctx.synchronize()
        tot = part[0] + part[1]
        torch.cuda.synchronize()
        ctx.finish_sum_dev(tot.data_ptr(), dcc.data_ptr(), nx, ny, -(L - 1), 0, NDIR, 1, "none", 0, ny, dout.data_ptr(),
                           dcost.data_ptr())
        ctx.synchronize()
        # another summation order: equal up to rounding (labels may flip at near-ties; costs within 1e-5 relative)
        o = O.orc_mgm(cc, None, -(L - 1), P1, P2, NDIR, K, felz, 1)
        fin = np.isfinite(o["outcost"])
        assert np.allclose(dcost.cpu().numpy()[fin], o["outcost"][fin], rtol=1e-5)
        assert np.mean(dout.cpu().numpy() != o["out"]) < 0.02
        ctx.sweeps_release()


@pytest.mark.parametrize("felz", [0, 1])
@pytest.mark.parametrize("weighted", [0, 1])
def test_mgm_16_sweeps(ctx, felz, weighted):
    """-O 16: sweeps 8-15 as defined by this build (knight-move neighbour chains; the reference is
    undefined there) against the oracle port's restatement of the same definition: every update variant, chained bands."""
    for K in [1, 2, 3, 4]:
        for (nx, ny, L) in [(41, 29, 12), (30, 52, 35)]:
            cc = synth_volume(nx, ny, L, seed=K + nx, real=True)
            w = synth_weights(nx, ny, seed=K) if weighted else None
            P1, P2 = ((8, 32) if K % 2 else (3.5, 21.25)) if not felz else ((2, 20000) if K != 4 else (1.5, 11))
            for rows in (0, 5):
                ctx.set_rows_per_band(rows)
                for NDIR in (16, 11):
                    r = ctx.mgm(cc, w, -(L - 1), P1, P2, NDIR, K, felz, 1)
                    o = O.orc_mgm(cc, w, -(L - 1), P1, P2, NDIR, K, felz, 1)
                    assert same(r["out"], o["out"]) and same(r["S"], o["S"]) and same(r["outcost"], o["outcost"]), \
                        (K, felz, weighted, nx, ny, L, rows, NDIR, mism(r["S"], o["S"]))
    ctx.set_rows_per_band(0)
    # the whole path with 16 sweeps
    u, v = synth_pair(97, 61, 20, seed=5)
    kw = dict(P1=8.0, P2=32.0, NDIR=16, distance="ncc", refinement="parabola")
    out, cost = ctx.stereo(u, v, dmin=-19, dmax=0, MGM=4 - felz, use_felzenszwalb_potentials=felz, census_ncc_win=3,
                           aP=4.0 if weighted else 1.0, aThresh=12.0, **kw)
    ref = O.orc_pipeline(u, v, -19, 0, K=4 - felz, felz=felz, win=3, aP=4.0 if weighted else 1.0, aThresh=12.0, **kw)
    assert same(out, ref["out"]) and same(cost, ref["outcost"])


def test_stereo_batch(ctx):
    """mgmb200_stereo_batch (images in, maps out, several pairs per aggregation launch) = one mgmb200_stereo per pair;
    pairs with and without image-dependent weights in one batch fall back to per-pair launches."""
    pairs = [synth_pair(83, 47, 20, seed=30 + i) for i in range(5)]
    us, vs = [p[0] for p in pairs], [p[1] for p in pairs]
    ctx.set_option("batch", 3)
    try:
        for kw in (dict(P1=8.0, P2=32.0, NDIR=8, MGM=4, distance="ad", refinement="vfit"),
                   dict(P1=2.0, P2=20000.0, NDIR=16, MGM=3, use_felzenszwalb_potentials=1, distance="census", refinement="cubic"),
                   dict(P1=8.0, P2=32.0, NDIR=8, MGM=2, distance="sd", aP=4.0, aThresh=12.0)):
            outs, costs = ctx.stereo_batch(us, vs, dmin=-19, dmax=0, **kw)
            for i in range(len(us)):
                o, c = ctx.stereo(us[i], vs[i], dmin=-19, dmax=0, **kw)
                assert same(outs[i], o) and same(costs[i], c), (kw, i)
    finally:
        ctx.set_option("reset")


def test_aggregate_batch(ctx):
    """mgmb200_aggregate_batch_dev: several pairs per launch (more pairs than the in-flight limit, so several launches)
    give the bits of one mgmb200_aggregate_dev call per pair."""
    import torch
    nx, ny, L, npairs = 70, 45, 24, 5
    VS = ctx.padded_labels(L)
    ctx.set_option("batch", 2)
    try:
        for (K, felz, P1, P2, NDIR) in [(4, 0, 8.0, 32.0, 8), (3, 1, 2.0, 20000.0, 8), (2, 0, 8.0, 32.0, 16)]:
            ccs = [synth_volume(nx, ny, L, seed=100 + b, real=True) for b in range(npairs)]
            dccs = []
            for cc in ccs:
                d = torch.empty((ny, nx, VS), device="cuda")
                ctx.pad_volume_dev(torch.from_numpy(cc).cuda().data_ptr(), d.data_ptr(), nx, ny, L)
                dccs.append(d)
            outs = [torch.empty((ny, nx), device="cuda") for _ in range(npairs)]
            costs = [torch.empty((ny, nx), device="cuda") for _ in range(npairs)]
            for rows in (0, 9):
                ctx.set_rows_per_band(rows)
                ctx.aggregate_batch_dev([d.data_ptr() for d in dccs], nx, ny, -(L - 1), 0, P1, P2, NDIR, K, felz, 1, "vfit",
                                        [o.data_ptr() for o in outs], [c.data_ptr() for c in costs])
                ctx.synchronize()
                for b in range(npairs):
                    o = O.orc_mgm(ccs[b], None, -(L - 1), P1, P2, NDIR, K, felz, 1)
                    ro, rc = O.orc_refine(o["S"], -(L - 1), o["out"], o["outcost"], "vfit")
                    assert same(outs[b].cpu().numpy(), ro) and same(costs[b].cpu().numpy(), rc), (K, felz, NDIR, rows, b)
    finally:
        ctx.set_option("reset")
        ctx.set_rows_per_band(0)


# ------------------------------------------------------------------------------------------ per-pixel ranges
def _ragged(nx, ny, emin, emax, seed):
    rng = np.random.default_rng(seed)
    lo = rng.integers(emin, emin + 6, (ny, nx)).astype(np.float32)
    hi = (lo + rng.integers(3, 9, (ny, nx))).clip(max=emax).astype(np.float32)
    return lo, hi


def test_ranges_costvolume_mgm_refine(ctx):
    """dminI / dmaxI images: cost volume, aggregation of the INF-masked dense volume, WTA over a
    second set of ranges (as from the second TSGM_ITER iteration on), refinement -- bit-exact against the oracle"""
    import mgm_b200
    nx, ny, emin, emax = 67, 41, -20, 5
    u, v = synth_pair(nx, ny, 22, seed=3, nch=1)
    lo, hi = _ragged(nx, ny, emin, emax, 1)
    for dist in ["ad", "census", "ncc"]:
        assert same(ctx.costvolume_ranges(u, v, lo, hi, emin, emax, "none", dist, np.inf, 3),
                    O.orc_costvolume_ranges(u, v, lo, hi, emin, emax, "none", dist, np.inf, 3)), dist
    cc = O.orc_costvolume_ranges(u, v, lo, hi, emin, emax, "none", "ad", np.inf, 3)
    srs = [(lo, hi), (lo + 1, np.maximum(hi - 1, lo + 1)), (np.maximum(lo - 2, emin), np.minimum(hi + 2, emax))]
    # every update variant; truncated linear: (2,1,0) folds the out-of-range labels back in (mgm_core.cc:166-219),
    # the others convolve inside the receiving pixel's range (:229-281, consumer-side kernels with masked loads)
    cases = [(K, felz, wt) for K in (1, 2, 3, 4) for felz in (0, 1) for wt in (0, 1)]
    for (K, felz, weighted), fix in itertools.product(cases, [0, 1]):
        w = synth_weights(nx, ny, seed=K) if weighted else None
        P1, P2 = (8, 32) if not felz else (2, 20000)
        for slo, shi in srs:
            r = ctx.mgm_ranges(cc, lo, hi, w, emin, slo, shi, P1, P2, 8, K, felz, fix)
            o = O.orc_mgm_ranges(cc, lo, hi, w, emin, slo, shi, P1, P2, 8, K, felz, fix)
            tag = (K, felz, weighted, fix, mism(r["S"], o["S"]), mism(r["out"], o["out"]))
            assert same(r["S"], o["S"]), tag
            assert same(r["out"], o["out"]) and same(r["outcost"], o["outcost"]), tag
            a = ctx.subpixel_refinement_sgm_ranges(o["S"], slo, shi, emin, o["out"], o["outcost"], "vfit")
            b = O.orc_refine_ranges(o["S"], slo, shi, emin, o["out"], o["outcost"], "vfit")
            assert same(a[0], b[0]) and same(a[1], b[1]), tag
    # uniform ranges through the same entry point == the plain call
    full_lo, full_hi = np.full((ny, nx), emin, np.float32), np.full((ny, nx), emax, np.float32)
    ccu = O.orc_costvolume(u, v, emin, emax, "none", "ad", np.inf, 3)
    r = ctx.mgm_ranges(ccu, full_lo, full_hi, None, emin, full_lo, full_hi, 2, 20000, 8, 3, 1, 1)
    o = O.orc_mgm(ccu, None, emin, 2, 20000, 8, 3, 1, 1)
    assert same(r["S"], o["S"]) and same(r["out"], o["out"])
    with pytest.raises(mgm_b200.MgmError):
        ctx.mgm_ranges(cc, hi, lo, None, emin, lo, hi, 8, 32, 8, 2, 0, 1)   # empty ranges


# ------------------------------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("path", golden_files("pipeline"))
def test_golden_pipeline(ctx, path):
    g = load_golden(path)
    p = dict(g["params"])
    kw = dict(dmin=p.pop("dmin"), dmax=p.pop("dmax"), MGM=p.pop("K"), use_felzenszwalb_potentials=p.pop("felz", 0),
              census_ncc_win=p.pop("win", 3))
    kw.update(p)
    out, cost = ctx.stereo(g["u"], g["v"], **kw)
    assert same(out, g["out"]), (g["name"], mism(out, g["out"]))
    assert same(cost, g["outcost"]), g["name"]


@pytest.mark.parametrize("path", golden_files("volume"))
def test_golden_volume(ctx, path):
    g = load_golden(path)
    p = g["params"]
    r = ctx.mgm(g["cc"], g["w"], p["dmin"], p["P1"], p["P2"], p["NDIR"], p["K"], p["felz"], p["fix"])
    assert same(r["out"], g["out"]) and same(r["outcost"], g["outcost"])
    assert same(r["S"][r["S"].shape[0] // 2], g["S_row"])


@pytest.mark.parametrize("path", golden_files("flow_lr"))
def test_golden_flow_lr(ctx, path):
    """mgmb200_stereo_lr against the output files of the reference binary (both directions, median, left-right tests,
    -l map, back-projection; tests/golden/make_golden_flows.py)"""
    from tests.flow_checks import check_flow_lr
    check_flow_lr(load_golden(path), ctx)


@pytest.mark.parametrize("path", golden_files("flow_ranges"))
def test_golden_flow_ranges(ctx, path):
    """mgmb200_stereo_ranges against the output files of the reference binary (-m/-M range images, TSGM_ITER)"""
    from tests.flow_checks import check_flow_ranges
    check_flow_ranges(load_golden(path), ctx)


# ------------------------------------------------------------------------------------------ whole hot path
@pytest.mark.parametrize("kw", [
    dict(distance="census", census_ncc_win=5, NDIR=8, MGM=2, refinement="vfit"),                       # BASELINE cfg 2 shape of flags
    dict(distance="census", census_ncc_win=3, NDIR=8, MGM=3, use_felzenszwalb_potentials=1, P1=2, P2=20000,
         refinement="vfit"),                                                                          # cfg 3 / Makefile:17
    dict(distance="ad", NDIR=4, MGM=4, aP=4.0, aThresh=6.0, refinement="cubic"),                       # weights on
    dict(distance="ncc", census_ncc_win=5, NDIR=8, MGM=4, refinement="parabola"),                      # cfg 5 flags (8 sweeps)
])
def test_stereo_matches_oracle(ctx, kw):
    u, v = synth_pair(197, 75, 48, seed=1, nch=1)
    out, cost = ctx.stereo(u, v, dmin=-47, dmax=0, **kw)
    kk = dict(kw)
    kk["win"] = kk.pop("census_ncc_win", 3); kk["K"] = kk.pop("MGM"); kk["felz"] = kk.pop("use_felzenszwalb_potentials", 0)
    b = O.orc_pipeline(u, v, -47, 0, **kk)
    assert same(out, b["out"]), mism(out, b["out"])
    assert same(cost, b["outcost"])


# ------------------------------------------------------------------------------------------ full-size properties
@pytest.mark.parametrize("cfg", [dict(W=2048, H=1536, L=256, dist="census", win=3, K=3, felz=1, P1=2.0, P2=20000.0),
                                 dict(W=1920, H=1080, L=128, dist="census", win=5, K=2, felz=0, P1=8.0, P2=32.0),
                                 dict(W=1242, H=375, L=192, dist="census", win=5, K=4, felz=0, P1=8.0, P2=32.0),
                                 dict(W=4096, H=4096, L=64, dist="ncc", win=5, K=2, felz=0, P1=8.0, P2=32.0)])
def test_full_size_finish_fused_vs_separate(ctx, cfg):
    """BASELINE.json configs 2-5 shapes: the finish stage run as tiles inside the aggregation launch (completion flags,
    release/acquire across CTAs, thousands of tiles racing with the running bands) gives the bits of the separate
    finish kernel, sub-pixel refinement included, and is reproducible from run to run."""
    from bench import synth_pair as bench_pair
    W, H, L = cfg["W"], cfg["H"], cfg["L"]
    u, v = bench_pair(W, H, L, 1)
    kw = dict(dmin=-(L - 1), dmax=0, P1=cfg["P1"], P2=cfg["P2"], MGM=cfg["K"], NDIR=8, refinement="vfit",
              use_felzenszwalb_potentials=cfg["felz"], distance=cfg["dist"], census_ncc_win=cfg["win"])
    ctx.set_option("fused_finish", 1)   # also where the default would not fuse (short label vectors)
    runs = [ctx.stereo(u, v, **kw) for _ in range(3)]
    assert ctx.last_launch_info()["kernel_launches"] == 1
    ctx.set_option("no_fused_finish", 1)
    try:
        ref = ctx.stereo(u, v, **kw)
        assert ctx.last_launch_info()["kernel_launches"] == 2
    finally:
        ctx.set_option("reset")
    for out, cost in runs:
        assert same(out, ref[0]) and same(cost, ref[1]), (mism(out, ref[0]), mism(cost, ref[1]))
    assert (ref[0] != np.round(ref[0])).any()   # sub-pixel offsets present (vfit may give NaN on flat minima, like the reference)


@pytest.mark.parametrize("cfg", [dict(W=2048, H=1536, L=256, win=3, K=3, felz=1, P1=2.0, P2=20000.0),
                                 dict(W=1920, H=1080, L=128, win=5, K=2, felz=0, P1=8.0, P2=32.0)])
def test_full_size_properties(ctx, cfg):
    """BASELINE.json configs 3 and 2 at full size.  Size-independent properties of the exact algorithm:
    (1) the result does not depend on how the image is cut into bands (rows per band = default vs 17),
        which exercises the boundary hand-off between thread blocks tens of thousands of times;
    (2) a crop of rows/columns that starts at the scan origin of sweep 0 reproduces exactly, for NDIR=1,
        the corresponding part of the full-size result (messages only flow along the sweep);
    (3) WTA labels are integers inside the disparity range."""
    from bench import synth_pair as bench_pair
    W, H, L = cfg["W"], cfg["H"], cfg["L"]
    u, v = bench_pair(W, H, L, 0)
    base = dict(dmin=-(L - 1), dmax=0, P1=cfg["P1"], P2=cfg["P2"], MGM=cfg["K"],
                use_felzenszwalb_potentials=cfg["felz"], distance="census", census_ncc_win=cfg["win"])
    ctx.set_rows_per_band(0)
    out_a, cost_a = ctx.stereo(u, v, NDIR=8, refinement="none", **base)
    ctx.set_rows_per_band(17)
    out_b, cost_b = ctx.stereo(u, v, NDIR=8, refinement="none", **base)
    ctx.set_rows_per_band(0)
    assert same(out_a, out_b) and same(cost_a, cost_b)
    assert out_a.min() >= -(L - 1) and out_a.max() <= 0 and np.all(out_a == np.round(out_a))
    # (2) sweep 0 only: crop invariance, checked against the ORACLE on the crop (small enough for the CPU)
    out1, cost1 = ctx.stereo(u, v, NDIR=1, refinement="none", **base)
    cw, ch = 200, 40
    kk = dict(P1=cfg["P1"], P2=cfg["P2"], NDIR=1, K=cfg["K"], felz=cfg["felz"], distance="census", win=cfg["win"])
    # the census transform and the matching range need the right image to extend L-1 pixels to the left:
    # crop at the left border so the INF wedge is identical; rows/cols beyond the crop only influence
    # later scan positions of sweep 0 (it runs left->right, top->bottom), except the census window at the
    # crop's last row/column and the border rule there -> compare the interior of the crop.
    o = O.orc_pipeline(u[:ch, :cw], v[:ch, :cw], -(L - 1), 0, **kk)
    assert same(out1[:ch - 2, :cw - 2], o["out"][:ch - 2, :cw - 2])
    assert same(cost1[:ch - 2, :cw - 2], o["outcost"][:ch - 2, :cw - 2])


@pytest.mark.parametrize("cfg", [dict(W=640, H=480, L=256, win=3, K=3, felz=1, P1=2.0, P2=20000.0),      # headline kernel: nj=8 register path
                                 dict(W=640, H=480, L=128, win=5, K=2, felz=0, P1=8.0, P2=32.0),        # configs[1] kernel
                                 dict(W=600, H=375, L=192, win=3, K=4, felz=0, P1=8.0, P2=32.0, dist="ad"),   # configs[3] kernel
                                 dict(W=512, H=512, L=64, win=5, K=2, felz=0, P1=8.0, P2=32.0, dist="ncc")])  # 4 lanes per worker
def test_midsize_against_the_reference(ctx, cfg):
    """The label counts and update variants of BASELINE.json's configurations on mid-size frames, DEFAULT bands (several
    chained bands per sweep, sheared diagonal sweeps, fused finish tiles), all eight sweeps, compared bit for bit with
    the UNMODIFIED reference run here (oracle/_ref, mgm_core.cc:408-613): cost volume, aggregated volume S, WTA labels,
    costs, sub-pixel disparities."""
    if not O.have_ref():
        pytest.skip("oracle/_ref not built on this box")
    from bench import synth_pair as bench_pair
    W, H, L = cfg["W"], cfg["H"], cfg["L"]
    dist = cfg.get("dist", "census")
    pf = "census" if dist == "census" else "none"
    u, v = bench_pair(W, H, L, 3)
    fl = "_flat" if O.ref_lib("_flat") is not None else ""
    cc_ref = O.ref_costvolume(u, v, -(L - 1), 0, pf, dist, np.inf, cfg["win"], flavour=fl)
    cc = ctx.allocate_and_fill_sgm_costvolume(u, v, -(L - 1), 0, pf, dist, np.inf, cfg["win"])
    assert same(cc, cc_ref), mism(cc, cc_ref)
    r = O.ref_mgm(cc_ref, None, -(L - 1), cfg["P1"], cfg["P2"], 8, cfg["K"], cfg["felz"], 1, flavour=fl)
    ro, rc = O.ref_refine(r["S"], -(L - 1), r["out"], r["outcost"], "vfit", flavour=fl)
    g = ctx.mgm(cc, None, -(L - 1), cfg["P1"], cfg["P2"], 8, cfg["K"], cfg["felz"], 1)
    assert ctx.last_launch_info()["rows_axis"] >= 28    # default bands (small SGM frames take 40 or 28 rows per band)
    assert same(g["out"], r["out"]), mism(g["out"], r["out"])
    assert same(g["S"], r["S"]), mism(g["S"], r["S"])
    assert same(g["outcost"], r["outcost"])
    kw = dict(dmin=-(L - 1), dmax=0, P1=cfg["P1"], P2=cfg["P2"], MGM=cfg["K"], NDIR=8, refinement="vfit",
              use_felzenszwalb_potentials=cfg["felz"], distance=dist, census_ncc_win=cfg["win"])
    ctx.set_option("fused_finish", 1)
    try:
        out, cost = ctx.stereo(u, v, **kw)                  # the fused path: one launch, finish tiles
        assert ctx.last_launch_info()["kernel_launches"] == 1
    finally:
        ctx.set_option("reset")
    assert same(out, ro) and same(cost, rc), (mism(out, ro), mism(cost, rc))


# ------------------------------------------------------------------------------------------ command lines
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MGM = os.path.join(ROOT, "oracle", "_ref", "mgm")
REF_MGM_O = os.path.join(ROOT, "oracle", "_ref", "mgm_o")
OUR_MGM = os.path.join(ROOT, "mgm_b200", "bin", "mgm")
OUR_MGM_O = os.path.join(ROOT, "mgm_b200", "bin", "mgm_o")


# ------------------------------------------------------------------------------------------ N1/N2 post-processing
def _disparity_maps(nx, ny, seed):
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


@pytest.mark.parametrize("shape", [(37, 23), (300, 41), (3, 3), (1, 7), (64, 1)])
def test_postprocessing_stages(ctx, shape):
    """leftright_test (mgm.cc:68-91), median_filter (img_tools.h:203-238), update_dmin_dmax (mgm.cc:120-158),
    back-projection (mgm.cc:432-443) against the oracle, NaN / INF / out-of-image disparities included"""
    nx, ny = shape
    dl, dr = _disparity_maps(nx, ny, nx + ny)
    for tau in (1.0, 0.0, 2.5):
        assert same(ctx.leftright_test(dl, dr, tau), O.orc_leftright(dl, dr, tau))
        assert same(ctx.leftright_test(dr, dl, tau), O.orc_leftright(dr, dl, tau))
    for r in (0, 1, 2, 3, 7):
        assert same(ctx.median_filter(dl, r), O.orc_median(dl, r)), r
    rgb = np.stack([dl, dr, dl + 1])
    assert same(ctx.median_filter(rgb, 1), O.orc_median(rgb, 1))
    lo = np.full((ny, nx), -12, np.float32)
    hi = np.full((ny, nx), 3, np.float32)
    for off, kw in ((dl, {}), (np.full((ny, nx), np.nan, np.float32), {}), (dl, dict(slack=-2, radius=1)), (dr, dict(slack=0, radius=0))):
        a, b = ctx.update_dmin_dmax(off, lo, hi, **kw), O.orc_update_range(off, lo, hi, **kw)
        assert same(a[0], b[0]) and same(a[1], b[1]) and a[2] == b[2], kw
    for nch in (1, 3):
        u, v = synth_pair(nx, ny, 8, seed=2, nch=nch)
        off = np.where(np.isfinite(dl) & (np.abs(dl) < 1e6), dl, np.float32(np.nan))
        assert same(ctx.backproject(off, u, v), O.orc_backproject(off, u, v))
    with pytest.raises(Exception):
        ctx.median_filter(dl, 8)


@pytest.mark.parametrize("kw", [
    dict(distance="ad", MGM=2, NDIR=4, median=0, testlrrl=1),
    dict(distance="census", MGM=3, NDIR=8, use_felzenszwalb_potentials=1, P1=2.0, P2=20000.0, refinement="vfit", median=1, testlrrl=1),
    dict(distance="sd", MGM=4, NDIR=8, aP=4.0, aThresh=9.0, refinement="cubic", median=2, testlrrl=1, testlrrl_tau=2.0),
    dict(distance="ad", MGM=1, NDIR=8, median=1, testlrrl=0),
])
def test_stereo_lr_flow(ctx, kw):
    """mgmb200_stereo_lr = the CLI flow mgm.cc:372-443: L->R, median, R->L on the mirrored range, median, both
    left-right tests on the untested maps, back-projection"""
    kw = dict(kw)
    median, lr, tau = kw.pop("median"), kw.pop("testlrrl"), kw.pop("testlrrl_tau", 1.0)
    nch = 3 if kw["distance"] == "sd" else 1
    u, v = synth_pair(97, 43, 20, seed=9, nch=nch)
    r = ctx.stereo_lr(u, v, -19, 3, testlrrl=lr, testlrrl_tau=tau, median=median, want_backproj=True, **kw)
    okw = dict(P1=kw.get("P1", 8.0), P2=kw.get("P2", 32.0), NDIR=kw["NDIR"], K=kw["MGM"], felz=kw.get("use_felzenszwalb_potentials", 0),
               aP=kw.get("aP", 1.0), aThresh=kw.get("aThresh", 5.0), distance=kw["distance"], refinement=kw.get("refinement", "none"))
    L = O.orc_pipeline(u, v, -19, 3, **okw)
    offL = O.orc_median(L["out"], median) if median else L["out"]
    assert same(r["out_nolr"], offL)
    out = offL
    if lr:
        R = O.orc_pipeline(v, u, -3, 19, **okw)
        offR = O.orc_median(R["out"], median) if median else R["out"]
        out = O.orc_leftright(offL, offR, tau)
        assert same(r["outR"], O.orc_leftright(offR, offL, tau)) and same(r["outcostR"], R["outcost"])
        assert np.isnan(out).any() and np.isfinite(out).any()
    assert same(r["out"], out) and same(r["outcost"], L["outcost"])
    assert same(r["backproj"], O.orc_backproject(out, u, v))


@pytest.mark.parametrize("kw", [
    dict(distance="ad", MGM=2, NDIR=8, iters=3, ragged=True),
    dict(distance="census", MGM=3, NDIR=8, use_felzenszwalb_potentials=1, P1=2.0, P2=20000.0, refinement="vfit", iters=2, ragged=True),
    dict(distance="sd", MGM=4, NDIR=4, aP=4.0, aThresh=9.0, refinement="cubic", iters=2, ragged=False),
    dict(distance="ad", MGM=2, NDIR=8, use_felzenszwalb_potentials=1, P1=2.0, P2=20000.0, iters=1, ragged=True),
])
def test_stereo_ranges_flow(ctx, kw):
    """mgmb200_stereo_ranges = mgm.cc:372-395 on the device: cost volume over the range images once, then per
    TSGM_ITER iteration mgm + refinement over the current ranges and update_dmin_dmax + non-finite repair"""
    kw = dict(kw)
    iters, ragged = kw.pop("iters"), kw.pop("ragged")
    nch = 3 if kw["distance"] == "sd" else 1
    nx, ny = 83, 37
    u, v = synth_pair(nx, ny, 18, seed=11, nch=nch)
    if ragged:
        lo, hi = _ragged(nx, ny, -20, 5, 5)
    else:
        lo, hi = np.full((ny, nx), -17, np.float32), np.full((ny, nx), 2, np.float32)
    out, cost, lo2, hi2 = ctx.stereo_ranges(u, v, lo, hi, tsgm_iter=iters, **kw)
    # the oracle's composition, over an envelope wide enough for every iteration
    emin, emax = int(lo.min()) - 3 * (iters - 1), int(hi.max()) + 3 * (iters - 1)
    P1, P2 = np.float32(kw.get("P1", 8.0)) * nch, np.float32(kw.get("P2", 32.0)) * nch
    w = O.orc_weights(u, kw.get("aP", 1.0), kw.get("aThresh", 5.0))
    cc = O.orc_costvolume_ranges(u, v, lo, hi, emin, emax, "none", kw["distance"], np.inf, 3)
    slo, shi = lo.copy(), hi.copy()
    for _ in range(iters):
        r = O.orc_mgm_ranges(cc, lo, hi, w, emin, slo, shi, P1, P2, kw["NDIR"], kw["MGM"], kw.get("use_felzenszwalb_potentials", 0), 1)
        o, oc = O.orc_refine_ranges(r["S"], slo, shi, emin, r["out"], r["outcost"], kw.get("refinement", "none"))
        slo, shi, (gmin, gmax) = O.orc_update_range(o, slo, shi)
        slo[~np.isfinite(slo)] = gmin
        shi[~np.isfinite(shi)] = gmax
    assert same(out, o) and same(cost, oc), (mism(out, o), mism(cost, oc))
    assert same(lo2, slo) and same(hi2, shi)


def _pnm(path, a):
    a = np.clip(a, 0, 255).astype(np.uint8)
    if a.shape[0] == 1:
        open(path, "wb").write(b"P5\n%d %d\n255\n" % (a.shape[2], a.shape[1]) + a[0].tobytes())
    else:
        open(path, "wb").write(b"P6\n%d %d\n255\n" % (a.shape[2], a.shape[1]) + np.transpose(a, (1, 2, 0)).tobytes())


@pytest.mark.skipif(not (os.path.exists(REF_MGM) and os.path.exists(OUR_MGM)), reason="CLI binaries not built")
@pytest.mark.parametrize("case", [
    (dict(TSGM="2"), ["-r", "-23", "-R", "4", "-t", "ad", "-O", "4"], 1),                                   # BASELINE cfg 1 flags
    (dict(MEDIAN="1", CENSUS_NCC_WIN="3", USE_TRUNCATED_LINEAR_POTENTIALS="1", TSGM="3"),
     ["-P2", "20000", "-P1", "2", "-r", "-23", "-R", "4", "-t", "census", "-s", "vfit", "-O", "8"], 3),     # Makefile:17
    (dict(MEDIAN="1", USE_TRUNCATED_LINEAR_POTENTIALS="1", TSGM="3"),
     ["-P2", "20000", "-P1", "4", "-r", "-23", "-R", "4", "-p", "sobel_x", "-truncDist", "63", "-s", "vfit", "-O", "8"], 3),  # Makefile:18
    (dict(TESTLRRL="0"), ["-r", "-23", "-R", "4", "-aP2", "4", "-aThresh", "9", "-s", "cubic", "-O", "8"], 1),  # CLI defaults, weights
    (dict(TSGM_ITER="2", TSGM="2"), ["-r", "-23", "-R", "4", "-t", "census", "-s", "vfit", "-O", "8"], 1),        # range update between iterations
    (dict(TSGM_ITER="3", TSGM="3", USE_TRUNCATED_LINEAR_POTENTIALS="1", MEDIAN="1"),
     ["-P1", "2", "-P2", "20000", "-r", "-23", "-R", "4", "-t", "census", "-s", "vfit", "-O", "8"], 1),
    (dict(TSGM="4", RANGES="1"), ["-r", "-23", "-R", "4", "-t", "ad", "-s", "parabola", "-O", "4", "-p", "gblur"], 3),  # -m/-M range images
    (dict(TSGM="2", RANGES="1", USE_TRUNCATED_LINEAR_POTENTIALS="1", TSGM_ITER="2"),
     ["-P1", "2", "-P2", "20000", "-r", "-23", "-R", "4", "-t", "sd", "-O", "8"], 1),
    (dict(TSGM="4", RANGES="1", MGMB200_STEPWISE="1"), ["-r", "-23", "-R", "4", "-t", "ad", "-s", "parabola", "-O", "4"], 3),  # call-by-call mirror functions
    (dict(TSGM_ITER="2", TSGM="3", USE_TRUNCATED_LINEAR_POTENTIALS="1", MGMB200_STEPWISE="1"),
     ["-P1", "2", "-P2", "20000", "-r", "-23", "-R", "4", "-t", "census", "-s", "vfit", "-O", "8"], 1),
    (dict(TSGM="3", RANGES="1", USE_TRUNCATED_LINEAR_POTENTIALS="1", MEDIAN="1"),                             # windowed min-convolution
     ["-P1", "2", "-P2", "20000", "-r", "-23", "-R", "4", "-t", "census", "-s", "vfit", "-O", "8", "-aP2", "3"], 1),
])
def test_cli_matches_reference_cli(tmp_path, case):
    """the whole `mgm` command (both LR directions, median, LR test, back-projection, console output)"""
    env_extra, args, nch = case
    env_extra = dict(env_extra)
    u, v = synth_pair(120, 64, 24, seed=4, nch=nch)
    _pnm(str(tmp_path / "u.pnm"), u)
    _pnm(str(tmp_path / "v.pnm"), v)
    if env_extra.pop("RANGES", None):   # per-pixel range images around the true disparity (-m / -M, mgm.cc:342-353)
        rng = np.random.default_rng(7)
        lo = (-14 + rng.integers(-6, 3, (64, 120))).astype(np.float32)
        hi = (lo + rng.integers(0, 14, (64, 120))).astype(np.float32)   # some max < min+1: the CLI repairs them
        lo[5, 7] = np.nan
        np.save(str(tmp_path / "dmin.npy"), lo)
        np.save(str(tmp_path / "dmax.npy"), hi)
        args = args + ["-m", str(tmp_path / "dmin.npy"), "-M", str(tmp_path / "dmax.npy")]
    outs = {}
    for tag, exe in (("ref", REF_MGM), ("our", OUR_MGM)):
        names = [str(tmp_path / ("%s_%s.npy" % (tag, k))) for k in ("disp", "cost", "back")]
        env = dict(os.environ, **env_extra)
        r = subprocess.run([exe] + args + [str(tmp_path / "u.pnm"), str(tmp_path / "v.pnm")] + names + ["-l", str(tmp_path / (tag + "_nolr.npy"))],
                           env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs[tag] = (r.stdout, [np.squeeze(np.load(n)) for n in names + [str(tmp_path / (tag + "_nolr.npy"))]])
    assert outs["our"][0].replace(" USING IMAGE DEPENDENT WEIGHTS\n", "") == outs["ref"][0].replace(" USING IMAGE DEPENDENT WEIGHTS\n", "")
    for a, b, what in zip(outs["our"][1], outs["ref"][1], ("disp", "cost", "back", "nolr")):
        assert a.shape == b.shape and same(a, b), (what, mism(a, b))


@pytest.mark.skipif(not (os.path.exists(REF_MGM_O) and os.path.exists(OUR_MGM_O)), reason="mgm_o binaries not built")
@pytest.mark.parametrize("argv", [["8", "32", "2", "0"], ["2", "20000", "4", "1"], ["8", "32", "1", "0"]])
def test_mgm_o_file_protocol(tmp_path, argv):
    """matlab/mgm_o.cc: input.bin -> output.bin"""
    ncol, nrow, nlab, NDIR = 47, 31, 12, 8
    rng = np.random.default_rng(5)
    costs = rng.integers(0, 64, (nlab, nrow, ncol)).astype(np.float32)
    w = np.where(rng.random((8, nrow, ncol)) < 0.25, 3.0, 1.0).astype(np.float32)
    with open(tmp_path / "input.bin", "wb") as f:
        f.write(np.array([ncol, nrow, nlab, NDIR], np.int32).tobytes() + costs.tobytes() + w.tobytes())
    res = {}
    for tag, exe in (("ref", REF_MGM_O), ("our", OUR_MGM_O)):
        out = tmp_path / (tag + ".bin")
        r = subprocess.run([exe, str(tmp_path / "input.bin"), str(out)] + argv, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        res[tag] = np.fromfile(out, np.float32)
    assert same(res["our"], res["ref"])
