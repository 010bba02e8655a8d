"""TEST INFRASTRUCTURE ONLY.

ctypes bindings for the two CPU checkers:

* ``orc_*``  -- oracle/liboracle.so, our plain-C restatement (mgm_oracle.c)
* ``ref_*``  -- oracle/_ref/libmgmref.so, the unmodified reference compiled from
  /root/reference by oracle/Makefile (travels to the GPU box as a prebuilt file)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product package
``mgm_b200`` never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_c_float_p = ctypes.POINTER(ctypes.c_float)
_c_double_p = ctypes.POINTER(ctypes.c_double)

DISTANCES = ["ad", "sd", "census", "ncc", "btad", "btsd"]
PREFILTERS = ["none", "census", "sobelx", "gblur"]
REFINEMENTS = ["none", "vfit", "parabola", "cubic", "parabolaOCV"]


def build(force=False):
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref."""
    args = ["make", "-C", _HERE, "-s"]
    if force:
        subprocess.check_call(args + ["clean"])
    subprocess.check_call(args + ["all"])


def _load(path):
    return ctypes.CDLL(path) if os.path.exists(path) else None


_orc = None
_ref = {}


def orc_lib():
    global _orc
    if _orc is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _orc = ctypes.CDLL(path)
    return _orc


def ref_lib(flavour=""):
    """flavour: '' (OpenMP, std::vector<Dvec>), '_serial', '_flat' (DVEC_ALLOCATION_HACK)."""
    if flavour not in _ref:
        _ref[flavour] = _load(os.path.join(_HERE, "_ref", "libmgmref%s.so" % flavour))
    return _ref[flavour]


def have_ref():
    return ref_lib() is not None


def _fp(a):
    return None if a is None else a.ctypes.data_as(_c_float_p)


def _img(a):
    """accepts (H,W) or (C,H,W) float arrays -> contiguous planar float32, nx, ny, nch"""
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim == 2:
        a = a[None]
    nch, ny, nx = a.shape
    return a, nx, ny, nch


# ----------------------------------------------------------------------------- oracle (C port)
def orc_weights(u, aP, aThresh):
    u, nx, ny, nch = _img(u)
    w = np.empty((8, ny, nx), np.float32)
    orc_lib().orc_weights(_fp(u), nx, ny, nch, ctypes.c_float(aP), ctypes.c_float(aThresh), _fp(w))
    return w


def orc_census(u, win):
    u, nx, ny, nch = _img(u)
    lib = orc_lib()
    nw = lib.orc_census_nwords(nch, win)
    out = np.empty((nw, ny, nx), np.uint32)
    lib.orc_census(_fp(u), nx, ny, nch, win, out.ctypes.data_as(ctypes.c_void_p))
    return out


def orc_costvolume(u, v, dmin, dmax, prefilter="none", distance="ad", truncDist=np.inf, win=3):
    u, nx, ny, nch = _img(u)
    v, vnx, vny, vnch = _img(v)
    assert nch == vnch
    lib = orc_lib()
    pf = lib.orc_prefilter_index(prefilter.encode())
    di = lib.orc_distance_index(distance.encode())
    # consistency fix of mgm_costvolume.h:358-362: either one selects the census prefilter; the cost FUNCTION was picked
    # before (:355), so "-p census -t ad" runs AD on the census bit strings (modelled by the C port)
    if di == 2:
        pf = 1
    L = dmax - dmin + 1
    cc = np.empty((ny, nx, L), np.float32)
    rc = lib.orc_costvolume(_fp(u), _fp(v), nx, ny, nch, vnx, vny, dmin, dmax, pf, di,
                            ctypes.c_float(truncDist), win, _fp(cc))
    if rc != 0:
        raise ValueError("orc_costvolume: unsupported combination (%d)" % rc)
    return cc


def orc_mgm(cc, w, dmin, P1, P2, NDIR, K, felz=0, fix=1, want_S=True, want_passes=False):
    cc = np.ascontiguousarray(cc, np.float32)
    ny, nx, L = cc.shape
    if w is not None:
        w = np.ascontiguousarray(w, np.float32)
        assert w.shape == (8, ny, nx)
    out = np.empty((ny, nx), np.float32)
    outcost = np.empty((ny, nx), np.float32)
    S = np.empty_like(cc) if want_S else None
    Lp = np.empty((NDIR,) + cc.shape, np.float32) if want_passes else None
    weighted = orc_lib().orc_mgm(_fp(cc), _fp(w), nx, ny, L, dmin, ctypes.c_float(P1), ctypes.c_float(P2),
                                 NDIR, K, felz, fix, _fp(out), _fp(outcost), _fp(S), _fp(Lp))
    res = dict(out=out, outcost=outcost, S=S, weighted=bool(weighted))
    if want_passes:
        res["passes"] = Lp
    return res


def orc_refine(S, dmin, out, outcost, refinement="none"):
    S = np.ascontiguousarray(S, np.float32)
    ny, nx, L = S.shape
    out = np.array(out, np.float32, copy=True)
    outcost = np.array(outcost, np.float32, copy=True)
    lib = orc_lib()
    m = lib.orc_refinement_index(refinement.encode())
    lib.orc_refine(_fp(S), nx, ny, dmin, dmin + L - 1, _fp(out), _fp(outcost), m)
    return out, outcost


def orc_pipeline(u, v, dmin, dmax, P1=8.0, P2=32.0, NDIR=4, K=4, felz=0, fix=1, aP=1.0, aThresh=5.0,
                 prefilter="none", distance="ad", truncDist=np.inf, win=3, refinement="none"):
    """mgm.cc:356-385 for one direction of the LR pair (P1,P2 scaled by nch like mgm.cc:356-357)."""
    uu, nx, ny, nch = _img(u)
    w = orc_weights(u, aP, aThresh)
    cc = orc_costvolume(u, v, dmin, dmax, prefilter, distance, truncDist, win)
    r = orc_mgm(cc, w, dmin, np.float32(P1) * nch, np.float32(P2) * nch, NDIR, K, felz, fix)
    out, outcost = orc_refine(r["S"], dmin, r["out"], r["outcost"], refinement)
    return dict(out=out, outcost=outcost, S=r["S"], cc=cc, w=w, wta=r["out"])


def _ranges(lo, hi, ny, nx):
    lo = np.ascontiguousarray(np.broadcast_to(np.asarray(lo, np.float32), (ny, nx)))
    hi = np.ascontiguousarray(np.broadcast_to(np.asarray(hi, np.float32), (ny, nx)))
    return lo, hi


def orc_costvolume_ranges(u, v, dminI, dmaxI, emin, emax, prefilter="none", distance="ad", truncDist=np.inf, win=3):
    """per-pixel ranges: dense (H,W,emax-emin+1) volume, +INF outside [dminI,dmaxI]"""
    u, nx, ny, nch = _img(u)
    v, vnx, vny, vnch = _img(v)
    lib = orc_lib()
    pf = lib.orc_prefilter_index(prefilter.encode())
    di = lib.orc_distance_index(distance.encode())
    if di == 2:
        pf = 1
    lo, hi = _ranges(dminI, dmaxI, ny, nx)
    cc = np.empty((ny, nx, emax - emin + 1), np.float32)
    rc = lib.orc_costvolume_ranges(_fp(u), _fp(v), nx, ny, nch, vnx, vny, _fp(lo), _fp(hi), emin, emax, pf, di,
                                   ctypes.c_float(truncDist), win, _fp(cc))
    if rc != 0:
        raise ValueError("orc_costvolume_ranges: unsupported combination (%d)" % rc)
    return cc


def orc_mgm_ranges(cc, ccmin, ccmax, w, emin, smin, smax, P1, P2, NDIR, K, felz=0, fix=1):
    cc = np.ascontiguousarray(cc, np.float32)
    ny, nx, L = cc.shape
    clo, chi = _ranges(ccmin, ccmax, ny, nx)
    slo, shi = _ranges(smin, smax, ny, nx)
    if w is not None:
        w = np.ascontiguousarray(w, np.float32)
    out = np.empty((ny, nx), np.float32)
    outcost = np.empty((ny, nx), np.float32)
    S = np.empty_like(cc)
    orc_lib().orc_mgm_ranges(_fp(cc), _fp(clo), _fp(chi), _fp(w), nx, ny, L, emin, _fp(slo), _fp(shi),
                             ctypes.c_float(P1), ctypes.c_float(P2), NDIR, K, felz, fix, _fp(out), _fp(outcost), _fp(S))
    return dict(out=out, outcost=outcost, S=S)


def orc_refine_ranges(S, smin, smax, emin, out, outcost, refinement="none"):
    S = np.ascontiguousarray(S, np.float32)
    ny, nx, L = S.shape
    slo, shi = _ranges(smin, smax, ny, nx)
    out = np.array(out, np.float32, copy=True)
    outcost = np.array(outcost, np.float32, copy=True)
    lib = orc_lib()
    lib.orc_refine_ranges(_fp(S), _fp(slo), _fp(shi), nx, ny, emin, L, _fp(out), _fp(outcost),
                          lib.orc_refinement_index(refinement.encode()))
    return out, outcost


def orc_leftright(dx, Rdx, threshold=1.0):
    dx = np.array(dx, np.float32, copy=True)
    Rdx = np.ascontiguousarray(Rdx, np.float32)
    orc_lib().orc_leftright(_fp(dx), _fp(Rdx), dx.shape[1], dx.shape[0], Rdx.shape[1], ctypes.c_float(threshold))
    return dx


def orc_median(u, radius):
    u = np.ascontiguousarray(u, np.float32)
    out = np.empty_like(u)
    for a, o in zip(u.reshape(-1, *u.shape[-2:]), out.reshape(-1, *u.shape[-2:])):
        orc_lib().orc_median(_fp(a), a.shape[1], a.shape[0], int(radius), _fp(o))
    return out


def orc_update_range(off, lo, hi, slack=3, radius=2):
    off = np.ascontiguousarray(off, np.float32)
    lo = np.array(lo, np.float32, copy=True)
    hi = np.array(hi, np.float32, copy=True)
    mm = np.zeros(2, np.float32)
    orc_lib().orc_update_range(_fp(off), off.shape[1], off.shape[0], _fp(lo), _fp(hi), int(slack), int(radius), _fp(mm))
    return lo, hi, (float(mm[0]), float(mm[1]))


def orc_backproject(off, u, v):
    off = np.ascontiguousarray(off, np.float32)
    u, nx, ny, nch = _img(u)
    v, vnx, vny, _ = _img(v)
    syn = np.empty_like(u)
    orc_lib().orc_backproject(_fp(off), _fp(u), _fp(v), nx, ny, nch, vnx, vny, _fp(syn))
    return syn


def orc_scan_preds(p, nx=9, ny=7):
    o = (ctypes.c_int * 8)()
    orc_lib().orc_pass_scan_preds(p, nx, ny, o)
    return [(o[2 * k], o[2 * k + 1]) for k in range(4)]


# ----------------------------------------------------------------------------- the reference itself
def ref_weights(u, aP, aThresh, flavour=""):
    u, nx, ny, nch = _img(u)
    w = np.empty((8, ny, nx), np.float32)
    ref_lib(flavour).ref_weights(_fp(u), nx, ny, nch, ctypes.c_float(aP), ctypes.c_float(aThresh), _fp(w))
    return w


def ref_costvolume(u, v, dmin, dmax, prefilter="none", distance="ad", truncDist=np.inf, win=3, flavour=""):
    u, nx, ny, nch = _img(u)
    v, vnx, vny, vnch = _img(v)
    L = dmax - dmin + 1
    cc = np.empty((ny, nx, L), np.float32)
    ref_lib(flavour).ref_costvolume(_fp(u), _fp(v), nx, ny, nch, vnx, vny, dmin, dmax, prefilter.encode(),
                                    distance.encode(), ctypes.c_float(truncDist), win, _fp(cc))
    return cc


def ref_mgm(cc, w, dmin, P1, P2, NDIR, K, felz=0, fix=1, want_S=True, naive=0, flavour=""):
    cc = np.ascontiguousarray(cc, np.float32)
    ny, nx, L = cc.shape
    if w is not None:
        w = np.ascontiguousarray(w, np.float32)
    out = np.empty((ny, nx), np.float32)
    outcost = np.empty((ny, nx), np.float32)
    S = np.empty_like(cc) if want_S else None
    lib = ref_lib(flavour)
    lib.ref_mgm.restype = ctypes.c_double
    t = lib.ref_mgm(_fp(cc), _fp(w), nx, ny, dmin, dmin + L - 1, ctypes.c_float(P1), ctypes.c_float(P2), NDIR,
                    K, felz, fix, naive, _fp(out), _fp(outcost), _fp(S))
    return dict(out=out, outcost=outcost, S=S, seconds=t)


def ref_refine(S, dmin, out, outcost, refinement="none", flavour=""):
    S = np.ascontiguousarray(S, np.float32)
    ny, nx, L = S.shape
    out = np.array(out, np.float32, copy=True)
    outcost = np.array(outcost, np.float32, copy=True)
    ref_lib(flavour).ref_refine(_fp(S), nx, ny, dmin, dmin + L - 1, _fp(out), _fp(outcost), refinement.encode())
    return out, outcost


def ref_pipeline(u, v, dmin, dmax, P1=8.0, P2=32.0, NDIR=4, K=4, felz=0, fix=1, aP=1.0, aThresh=5.0,
                 prefilter="none", distance="ad", truncDist=np.inf, win=3, refinement="none", flavour=""):
    u, nx, ny, nch = _img(u)
    v, _, _, _ = _img(v)
    out = np.empty((ny, nx), np.float32)
    outcost = np.empty((ny, nx), np.float32)
    times = (ctypes.c_double * 4)()
    ref_lib(flavour).ref_pipeline(_fp(u), _fp(v), nx, ny, nch, dmin, dmax, ctypes.c_float(P1), ctypes.c_float(P2),
                                  NDIR, K, felz, fix, ctypes.c_float(aP), ctypes.c_float(aThresh),
                                  prefilter.encode(), distance.encode(), ctypes.c_float(truncDist), win,
                                  refinement.encode(), _fp(out), _fp(outcost), times)
    return dict(out=out, outcost=outcost, times=dict(weights=times[0], costvolume=times[1], mgm=times[2],
                                                     refine=times[3]))


def ref_costvolume_ranges(u, v, dminI, dmaxI, emin, emax, prefilter="none", distance="ad", truncDist=np.inf, win=3,
                          flavour=""):
    u, nx, ny, nch = _img(u)
    v, vnx, vny, vnch = _img(v)
    lo, hi = _ranges(dminI, dmaxI, ny, nx)
    cc = np.empty((ny, nx, emax - emin + 1), np.float32)
    ref_lib(flavour).ref_costvolume_ranges(_fp(u), _fp(v), nx, ny, nch, vnx, vny, _fp(lo), _fp(hi), emin, emax,
                                           prefilter.encode(), distance.encode(), ctypes.c_float(truncDist), win, _fp(cc))
    return cc


def ref_mgm_ranges(cc, ccmin, ccmax, w, emin, smin, smax, P1, P2, NDIR, K, felz=0, fix=1, flavour=""):
    cc = np.ascontiguousarray(cc, np.float32)
    ny, nx, L = cc.shape
    clo, chi = _ranges(ccmin, ccmax, ny, nx)
    slo, shi = _ranges(smin, smax, ny, nx)
    if w is not None:
        w = np.ascontiguousarray(w, np.float32)
    out = np.empty((ny, nx), np.float32)
    outcost = np.empty((ny, nx), np.float32)
    S = np.empty_like(cc)
    lib = ref_lib(flavour)
    lib.ref_mgm_ranges.restype = ctypes.c_double
    lib.ref_mgm_ranges(_fp(cc), _fp(clo), _fp(chi), _fp(w), nx, ny, emin, emin + L - 1, _fp(slo), _fp(shi),
                       ctypes.c_float(P1), ctypes.c_float(P2), NDIR, K, felz, fix, _fp(out), _fp(outcost), _fp(S))
    return dict(out=out, outcost=outcost, S=S)


def ref_refine_ranges(S, smin, smax, emin, out, outcost, refinement="none", flavour=""):
    S = np.ascontiguousarray(S, np.float32)
    ny, nx, L = S.shape
    slo, shi = _ranges(smin, smax, ny, nx)
    out = np.array(out, np.float32, copy=True)
    outcost = np.array(outcost, np.float32, copy=True)
    ref_lib(flavour).ref_refine_ranges(_fp(S), _fp(slo), _fp(shi), nx, ny, emin, emin + L - 1, _fp(out), _fp(outcost),
                                       refinement.encode())
    return out, outcost


# the reference's CLI translation unit as a library (oracle/ref_cli_harness.cc)
def refcli_lib():
    if "cli" not in _ref:
        _ref["cli"] = _load(os.path.join(_HERE, "_ref", "libmgmref_cli.so"))
    return _ref["cli"]


def ref_leftright(dx, Rdx, threshold=1.0):
    dx = np.array(dx, np.float32, copy=True)
    Rdx = np.ascontiguousarray(Rdx, np.float32)
    refcli_lib().refcli_leftright(_fp(dx), dx.shape[1], dx.shape[0], _fp(Rdx), Rdx.shape[1], Rdx.shape[0],
                                  ctypes.c_float(threshold))
    return dx


def ref_median(u, radius):
    a, nx, ny, nch = _img(u)
    out = np.empty_like(a)
    refcli_lib().refcli_median(_fp(a), nx, ny, nch, int(radius), _fp(out))
    return out.reshape(np.shape(u))


def ref_update_range(off, lo, hi, slack=3, radius=2):
    off = np.ascontiguousarray(off, np.float32)
    lo = np.array(lo, np.float32, copy=True)
    hi = np.array(hi, np.float32, copy=True)
    mm = np.zeros(2, np.float32)
    refcli_lib().refcli_update_dmin_dmax(_fp(off), off.shape[1], off.shape[0], _fp(lo), _fp(hi), int(slack), int(radius),
                                         _fp(mm))
    return lo, hi, (float(mm[0]), float(mm[1]))
