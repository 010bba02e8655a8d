"""ctypes binding of include/mgmb200.h and the host-side mirror of the reference API."""
import ctypes
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB = None

DISTANCES = ["ad", "sd", "census", "ncc", "btad", "btsd"]          # mgm_costvolume.h:170-183
PREFILTERS = ["none", "census", "sobelx", "gblur"]                # mgm_costvolume.h:194-200
REFINEMENTS = ["none", "vfit", "parabola", "cubic", "parabolaOCV"]  # mgm_refine.h:14-27

c_float_p = ctypes.POINTER(ctypes.c_float)


class MgmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mgmb200 error %d: %s" % (code, msg))
        self.code = code


def library_path():
    return os.environ.get("MGMB200_LIBRARY") or os.path.join(_HERE, "libmgmb200.so")   # override: development builds


def build_library(verbose=False):
    """Compile the CUDA sources for sm_100a into mgm_b200/libmgmb200.so (nvcc, in-tree)."""
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j4"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:])
        print(out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("building libmgmb200.so failed")
    return library_path()


def exported_symbols():
    """Function names declared in include/mgmb200.h."""
    text = open(os.path.join(_ROOT, "include", "mgmb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mgmb200_[a-z0-9_]+)\s*\(", text)))


class StereoParams(ctypes.Structure):
    """mgmb200_stereo_params: the CLI/env surface of mgm.cc:186-196,303-318."""
    _fields_ = [("dmin", ctypes.c_int), ("dmax", ctypes.c_int), ("P1", ctypes.c_float), ("P2", ctypes.c_float),
                ("NDIR", ctypes.c_int), ("MGM", ctypes.c_int), ("use_felzenszwalb_potentials", ctypes.c_int),
                ("sgm_fix_overcount", ctypes.c_int), ("aP", ctypes.c_float), ("aThresh", ctypes.c_float),
                ("prefilter", ctypes.c_char_p), ("distance", ctypes.c_char_p), ("truncDist", ctypes.c_float),
                ("census_ncc_win", ctypes.c_int), ("refinement", ctypes.c_char_p)]


class PostParams(ctypes.Structure):
    """mgmb200_post_params: TESTLRRL, TESTLRRL_TAU, MEDIAN (mgm.cc:194-196)."""
    _fields_ = [("testlrrl", ctypes.c_int), ("testlrrl_tau", ctypes.c_float), ("median", ctypes.c_int)]


def load_library():
    """Load libmgmb200.so; raises if it has not been built (no fallback of any kind)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` or "
                          "`make -C mgm_b200/csrc`" % path)
    lib = ctypes.CDLL(path)
    lib.mgmb200_last_error.restype = ctypes.c_char_p
    lib.mgmb200_volume_bytes.restype = ctypes.c_size_t
    for name in exported_symbols():
        getattr(lib, name)   # every declared entry point must be exported
    _LIB = lib
    return lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return None if a is None else a.ctypes.data_as(c_float_p)


def _img(a):
    a = _f32(a)
    if a.ndim == 2:
        a = a[None]
    if a.ndim != 3:
        raise ValueError("image must be (H,W) or (C,H,W)")
    nch, ny, nx = a.shape
    return a, nx, ny, nch


class Context:
    """One GPU context (mgmb200_create).  Methods mirror the reference functions on numpy arrays;
    ``*_dev`` methods take raw device pointers (ints, e.g. ``torch.Tensor.data_ptr()``)."""

    def __init__(self, device=-1):
        self.lib = load_library()
        self._ctx = ctypes.c_void_p()
        self._check(self.lib.mgmb200_create(int(device), ctypes.byref(self._ctx)))

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.mgmb200_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise MgmError(rc, self.lib.mgmb200_last_error().decode(errors="replace"))

    # ------------------------------------------------------------------ knobs
    def set_stream(self, cuda_stream):
        self._check(self.lib.mgmb200_set_stream(self._ctx, ctypes.c_void_p(int(cuda_stream) if cuda_stream else 0)))

    def set_rows_per_band(self, rows):
        self._check(self.lib.mgmb200_set_rows_per_band(self._ctx, int(rows)))

    def set_option(self, name, value=None):
        """Tuning / debugging knob (mgmb200_set_option); name "reset" re-reads the MGMB200_* environment."""
        self._check(self.lib.mgmb200_set_option(self._ctx, name.encode(), None if value is None else str(value).encode()))

    def synchronize(self):
        self._check(self.lib.mgmb200_synchronize(self._ctx))

    def last_launch_info(self):
        n, ra, rd, th = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        sm = ctypes.c_size_t()
        self._check(self.lib.mgmb200_last_launch_info(self._ctx, ctypes.byref(n), ctypes.byref(ra), ctypes.byref(rd),
                                                      ctypes.byref(th), ctypes.byref(sm)))
        return dict(kernel_launches=n.value, rows_axis=ra.value, rows_diag=rd.value, threads_per_cta=th.value,
                    smem_bytes=sm.value)

    # ------------------------------------------------------------------ reference operator interface (host arrays)
    def compute_mgm_weights(self, u, aP, aThresh):
        """compute_mgm_weights (mgm_weights.h:63) -> (8,H,W) planes W,E,S,N,NW,NE,SE,SW."""
        u, nx, ny, nch = _img(u)
        w = np.empty((8, ny, nx), np.float32)
        self._check(self.lib.mgmb200_compute_mgm_weights(self._ctx, _fp(u), nx, ny, nch, ctypes.c_float(aP),
                                                         ctypes.c_float(aThresh), _fp(w)))
        return w

    def allocate_and_fill_sgm_costvolume(self, u, v, dmin, dmax, prefilter="none", distance="ad",
                                         truncDist=np.inf, census_ncc_win=3):
        """allocate_and_fill_sgm_costvolume (mgm_costvolume.h:337) -> (H,W,L) float32."""
        u, nx, ny, nch = _img(u)
        v, vnx, vny, vnch = _img(v)
        if nch != vnch:
            raise ValueError("u and v must have the same number of channels")
        cc = np.empty((ny, nx, dmax - dmin + 1), np.float32)
        self._check(self.lib.mgmb200_costvolume(self._ctx, _fp(u), _fp(v), nx, ny, nch, vnx, vny, int(dmin), int(dmax),
                                                prefilter.encode(), distance.encode(), ctypes.c_float(truncDist),
                                                int(census_ncc_win), _fp(cc)))
        return cc

    def mgm(self, cc, w, dmin, P1, P2, NDIR, MGM, use_felzenszwalb_potentials=0, sgm_fix_overcount=1, want_S=True):
        """mgm (mgm_core.cc:408): cc (H,W,L), w (8,H,W) or None -> dict(out, outcost, S)."""
        cc = _f32(cc)
        ny, nx, L = cc.shape
        if w is not None:
            w = _f32(w)
            if w.shape != (8, ny, nx):
                raise ValueError("weights must be (8,H,W)")
        out = np.empty((ny, nx), np.float32)
        outcost = np.empty((ny, nx), np.float32)
        S = np.empty_like(cc) if want_S else None
        self._check(self.lib.mgmb200_mgm(self._ctx, _fp(cc), _fp(w), nx, ny, int(dmin), int(dmin) + L - 1,
                                         ctypes.c_float(P1), ctypes.c_float(P2), int(NDIR), int(MGM),
                                         int(use_felzenszwalb_potentials), int(sgm_fix_overcount), _fp(out),
                                         _fp(outcost), _fp(S)))
        return dict(out=out, outcost=outcost, S=S)

    # ------------------------------------------------------------------ per-pixel disparity ranges (dminI / dmaxI)
    @staticmethod
    def _range_imgs(lo, hi, ny, nx):
        lo = np.ascontiguousarray(np.broadcast_to(np.asarray(lo, np.float32), (ny, nx)))
        hi = np.ascontiguousarray(np.broadcast_to(np.asarray(hi, np.float32), (ny, nx)))
        return lo, hi

    def costvolume_ranges(self, u, v, dminI, dmaxI, emin, emax, prefilter="none", distance="ad", truncDist=np.inf,
                          census_ncc_win=3):
        """allocate_and_fill_sgm_costvolume with range images -> dense (H,W,emax-emin+1), +INF outside the ranges."""
        u, nx, ny, nch = _img(u)
        v, vnx, vny, vnch = _img(v)
        lo, hi = self._range_imgs(dminI, dmaxI, ny, nx)
        cc = np.empty((ny, nx, emax - emin + 1), np.float32)
        self._check(self.lib.mgmb200_costvolume_ranges(self._ctx, _fp(u), _fp(v), nx, ny, nch, vnx, vny, _fp(lo), _fp(hi),
                                                       int(emin), int(emax), prefilter.encode(), distance.encode(),
                                                       ctypes.c_float(truncDist), int(census_ncc_win), _fp(cc)))
        return cc

    def mgm_ranges(self, cc, ccmin, ccmax, w, emin, dminI, dmaxI, P1, P2, NDIR, MGM, use_felzenszwalb_potentials=0,
                   sgm_fix_overcount=1, want_S=True):
        """mgm() with per-pixel ranges: cc dense over [emin, emin+L-1] with vector ranges [ccmin,ccmax]; dminI/dmaxI
        are the ranges of the returned volume."""
        cc = _f32(cc)
        ny, nx, L = cc.shape
        clo, chi = self._range_imgs(ccmin, ccmax, ny, nx)
        slo, shi = self._range_imgs(dminI, dmaxI, ny, nx)
        if w is not None:
            w = _f32(w)
        out = np.empty((ny, nx), np.float32)
        outcost = np.empty((ny, nx), np.float32)
        S = np.empty_like(cc) if want_S else None
        self._check(self.lib.mgmb200_mgm_ranges(self._ctx, _fp(cc), _fp(clo), _fp(chi), _fp(w), nx, ny, int(emin),
                                                int(emin) + L - 1, _fp(slo), _fp(shi), ctypes.c_float(P1),
                                                ctypes.c_float(P2), int(NDIR), int(MGM), int(use_felzenszwalb_potentials),
                                                int(sgm_fix_overcount), _fp(out), _fp(outcost), _fp(S)))
        return dict(out=out, outcost=outcost, S=S)

    def subpixel_refinement_sgm_ranges(self, S, dminI, dmaxI, emin, out, outcost, refinement):
        S = _f32(S)
        ny, nx, L = S.shape
        slo, shi = self._range_imgs(dminI, dmaxI, ny, nx)
        out = np.array(out, np.float32, copy=True)
        outcost = np.array(outcost, np.float32, copy=True)
        self._check(self.lib.mgmb200_subpixel_refinement_sgm_ranges(self._ctx, _fp(S), _fp(slo), _fp(shi), nx, ny,
                                                                    int(emin), int(emin) + L - 1, _fp(out), _fp(outcost),
                                                                    refinement.encode()))
        return out, outcost

    def mgm_labelmajor(self, costs, w, P1, P2, NDIR, MGM, use_felzenszwalb_potentials=0):
        """matlab/mgm_o.cc protocol: costs (L,H,W) label-major planes -> labels (H,W)."""
        costs = _f32(costs)
        nlab, nrow, ncol = costs.shape
        if w is None:
            w = np.ones((8, nrow, ncol), np.float32)
        w = _f32(w)
        labels = np.empty((nrow, ncol), np.float32)
        outcost = np.empty((nrow, ncol), np.float32)
        self._check(self.lib.mgmb200_mgm_labelmajor(self._ctx, _fp(costs), _fp(w), ncol, nrow, nlab, ctypes.c_float(P1),
                                                    ctypes.c_float(P2), int(NDIR), int(MGM),
                                                    int(use_felzenszwalb_potentials), _fp(labels), _fp(outcost)))
        return labels, outcost

    def subpixel_refinement_sgm(self, S, dmin, out, outcost, refinement="none"):
        """subpixel_refinement_sgm (mgm_refine.h:40): returns refined copies of out, outcost."""
        S = _f32(S)
        ny, nx, L = S.shape
        out = np.array(out, np.float32, copy=True)
        outcost = np.array(outcost, np.float32, copy=True)
        self._check(self.lib.mgmb200_subpixel_refinement_sgm(self._ctx, _fp(S), nx, ny, int(dmin), int(dmin) + L - 1,
                                                             _fp(out), _fp(outcost), refinement.encode()))
        return out, outcost

    def stereo(self, u, v, dmin=-30, dmax=30, P1=8.0, P2=32.0, NDIR=4, MGM=4, use_felzenszwalb_potentials=0,
               sgm_fix_overcount=1, aP=1.0, aThresh=5.0, prefilter="none", distance="ad", truncDist=np.inf,
               census_ncc_win=3, refinement="none", out=None, outcost=None):
        """The hot path of mgm.cc:356-385 (one direction): images in, (disparity, cost) out.
        `out` / `outcost`: optional caller-owned float32 (ny, nx) arrays to receive the maps (e.g. pinned memory)."""
        u, nx, ny, nch = _img(u)
        v, vnx, vny, vnch = _img(v)
        if (vnx, vny, vnch) != (nx, ny, nch):
            raise ValueError("u and v must have the same shape")
        p = StereoParams(int(dmin), int(dmax), P1, P2, int(NDIR), int(MGM), int(use_felzenszwalb_potentials),
                         int(sgm_fix_overcount), aP, aThresh, prefilter.encode(), distance.encode(), truncDist,
                         int(census_ncc_win), refinement.encode())
        for name, a in (("out", out), ("outcost", outcost)):
            if a is not None and not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.shape == (ny, nx)
                                      and a.flags["C_CONTIGUOUS"]):
                raise ValueError("%s must be a C-contiguous float32 array of shape (%d, %d)" % (name, ny, nx))
        out = np.empty((ny, nx), np.float32) if out is None else out
        outcost = np.empty((ny, nx), np.float32) if outcost is None else outcost
        self._check(self.lib.mgmb200_stereo(self._ctx, _fp(u), _fp(v), nx, ny, nch, ctypes.byref(p), _fp(out),
                                            _fp(outcost)))
        return out, outcost

    def stereo_batch(self, us, vs, dmin=-30, dmax=30, P1=8.0, P2=32.0, NDIR=4, MGM=4, use_felzenszwalb_potentials=0,
                     sgm_fix_overcount=1, aP=1.0, aThresh=5.0, prefilter="none", distance="ad", truncDist=np.inf,
                     census_ncc_win=3, refinement="none", outs=None, outcosts=None):
        """mgmb200_stereo_batch: lists of images of one shape in, lists of (disparity, cost) maps out."""
        n = len(us)
        imgs = [(_img(a), _img(b)) for a, b in zip(us, vs)]
        (u0, nx, ny, nch) = imgs[0][0]
        for (a, ax, ay, ac), (b, bx, by, bc) in imgs:
            if (ax, ay, ac) != (nx, ny, nch) or (bx, by, bc) != (nx, ny, nch):
                raise ValueError("all images of a batch must have the same shape")
        p = StereoParams(int(dmin), int(dmax), P1, P2, int(NDIR), int(MGM), int(use_felzenszwalb_potentials),
                         int(sgm_fix_overcount), aP, aThresh, prefilter.encode(), distance.encode(), truncDist,
                         int(census_ncc_win), refinement.encode())
        outs = [np.empty((ny, nx), np.float32) for _ in range(n)] if outs is None else outs
        outcosts = [np.empty((ny, nx), np.float32) for _ in range(n)] if outcosts is None else outcosts
        arr = lambda xs: (c_float_p * n)(*[_fp(x) for x in xs])
        self._check(self.lib.mgmb200_stereo_batch(self._ctx, n, arr([i[0][0] for i in imgs]), arr([i[1][0] for i in imgs]), nx, ny,
                                                  nch, ctypes.byref(p), arr(outs), arr(outcosts)))
        return outs, outcosts

    # ------------------------------------------------------------------ post-processing of the CLI flow
    def leftright_test(self, dx, Rdx, threshold=1.0):
        """leftright_test (mgm.cc:68-91): returns the tested copy of dx."""
        dx = np.array(dx, np.float32, copy=True)
        Rdx = _f32(Rdx)
        (ny, nx), (rny, rnx) = dx.shape, Rdx.shape
        self._check(self.lib.mgmb200_leftright_test(self._ctx, _fp(dx), nx, ny, _fp(Rdx), rnx, rny,
                                                    ctypes.c_float(threshold)))
        return dx

    def median_filter(self, u, radius):
        """median_filter (img_tools.h:203-238) of an (H,W) or (C,H,W) image."""
        a, nx, ny, nch = _img(u)
        out = np.empty_like(a)
        self._check(self.lib.mgmb200_median_filter(self._ctx, _fp(a), nx, ny, nch, int(radius), _fp(out)))
        return out.reshape(np.shape(u))

    def update_dmin_dmax(self, outoff, dminI, dmaxI, slack=3, radius=2):
        """update_dmin_dmax (mgm.cc:120-158): returns (dminI, dmaxI, (gmin, gmax))."""
        off = _f32(outoff)
        ny, nx = off.shape
        lo = np.array(dminI, np.float32, copy=True).reshape(ny, nx)
        hi = np.array(dmaxI, np.float32, copy=True).reshape(ny, nx)
        g0, g1 = ctypes.c_float(), ctypes.c_float()
        self._check(self.lib.mgmb200_update_dmin_dmax(self._ctx, _fp(off), nx, ny, _fp(lo), _fp(hi), int(slack),
                                                      int(radius), ctypes.byref(g0), ctypes.byref(g1)))
        return lo, hi, (g0.value, g1.value)

    def backproject(self, outoff, u, v):
        """The back-projected image of mgm.cc:432-443."""
        off = _f32(outoff)
        u, nx, ny, nch = _img(u)
        v, vnx, vny, vnch = _img(v)
        if vnch != nch or off.shape != (ny, nx):
            raise ValueError("outoff must have the size of u, and v its channels")
        syn = np.empty_like(u)
        self._check(self.lib.mgmb200_backproject(self._ctx, _fp(off), _fp(u), _fp(v), nx, ny, nch, vnx, vny, _fp(syn)))
        return syn

    def stereo_lr(self, u, v, dmin=-30, dmax=30, P1=8.0, P2=32.0, NDIR=4, MGM=4, use_felzenszwalb_potentials=0,
                  sgm_fix_overcount=1, aP=1.0, aThresh=5.0, prefilter="none", distance="ad", truncDist=np.inf,
                  census_ncc_win=3, refinement="none", testlrrl=1, testlrrl_tau=1.0, median=0, want_backproj=False,
                  buffers=None):
        """The default command-line flow mgm.cc:372-443 on the device: both directions, median, left-right tests,
        back-projection.  Returns a dict with out, outcost, out_nolr and, with testlrrl, outR, outcostR.
        `buffers`: a dict returned by an earlier call with the same shapes, reused for the outputs."""
        u, nx, ny, nch = _img(u)
        v, vnx, vny, vnch = _img(v)
        if (vnx, vny, vnch) != (nx, ny, nch):
            raise ValueError("u and v must have the same shape")
        p = StereoParams(int(dmin), int(dmax), P1, P2, int(NDIR), int(MGM), int(use_felzenszwalb_potentials),
                         int(sgm_fix_overcount), aP, aThresh, prefilter.encode(), distance.encode(), truncDist,
                         int(census_ncc_win), refinement.encode())
        q = PostParams(int(testlrrl), testlrrl_tau, int(median))
        names = ["out", "outcost", "out_nolr"] + (["outR", "outcostR"] if testlrrl else [])
        r = {k: np.empty((ny, nx), np.float32) for k in names}
        if want_backproj:
            r["backproj"] = np.empty((nch, ny, nx), np.float32)
        if buffers is not None:
            for k in r:
                b = buffers.get(k)
                if not (isinstance(b, np.ndarray) and b.dtype == np.float32 and b.shape == r[k].shape and b.flags["C_CONTIGUOUS"]):
                    raise ValueError("buffers[%r] must be a C-contiguous float32 array of shape %s" % (k, r[k].shape))
                r[k] = b
        self._check(self.lib.mgmb200_stereo_lr(self._ctx, _fp(u), _fp(v), nx, ny, nch, ctypes.byref(p), ctypes.byref(q),
                                               _fp(r["out"]), _fp(r["outcost"]), _fp(r.get("outR")),
                                               _fp(r.get("outcostR")), _fp(r["out_nolr"]), _fp(r.get("backproj"))))
        return r

    def stereo_ranges(self, u, v, dminI, dmaxI, tsgm_iter=1, P1=8.0, P2=32.0, NDIR=4, MGM=4,
                      use_felzenszwalb_potentials=0, sgm_fix_overcount=1, aP=1.0, aThresh=5.0, prefilter="none",
                      distance="ad", truncDist=np.inf, census_ncc_win=3, refinement="none"):
        """One direction of mgm.cc:372-395 with range images and TSGM_ITER iterations, device resident.
        Returns (out, outcost, dminI, dmaxI) with the range images as updated by the last iteration."""
        u, nx, ny, nch = _img(u)
        v, vnx, vny, vnch = _img(v)
        if (vnx, vny, vnch) != (nx, ny, nch):
            raise ValueError("u and v must have the same shape")
        lo = np.array(np.broadcast_to(np.asarray(dminI, np.float32), (ny, nx)), np.float32, copy=True)
        hi = np.array(np.broadcast_to(np.asarray(dmaxI, np.float32), (ny, nx)), np.float32, copy=True)
        p = StereoParams(0, 0, P1, P2, int(NDIR), int(MGM), int(use_felzenszwalb_potentials), int(sgm_fix_overcount), aP,
                         aThresh, prefilter.encode(), distance.encode(), truncDist, int(census_ncc_win), refinement.encode())
        out, outcost = np.empty((ny, nx), np.float32), np.empty((ny, nx), np.float32)
        self._check(self.lib.mgmb200_stereo_ranges(self._ctx, _fp(u), _fp(v), nx, ny, nch, ctypes.byref(p), _fp(lo), _fp(hi),
                                                   int(tsgm_iter), _fp(out), _fp(outcost)))
        return out, outcost, lo, hi

    # ------------------------------------------------------------------ device-pointer interface
    @staticmethod
    def padded_labels(L):
        return (int(L) + 31) & ~31

    def weights_dev(self, d_u, nx, ny, nch, aP, aThresh, d_w):
        self._check(self.lib.mgmb200_weights_dev(self._ctx, ctypes.c_void_p(d_u), nx, ny, nch, ctypes.c_float(aP),
                                                 ctypes.c_float(aThresh), ctypes.c_void_p(d_w)))

    def costvolume_dev(self, d_u, d_v, nx, ny, nch, dmin, dmax, prefilter, distance, truncDist, win, d_cc,
                       vnx=None, vny=None):
        pf = self.lib.mgmb200_prefilter_index(prefilter.encode())
        di = self.lib.mgmb200_distance_index(distance.encode())
        if di == 2:
            pf = 1
        self._check(self.lib.mgmb200_costvolume_dev(self._ctx, ctypes.c_void_p(d_u), ctypes.c_void_p(d_v), nx, ny, nch,
                                                    vnx or nx, vny or ny, int(dmin), int(dmax), pf, di,
                                                    ctypes.c_float(truncDist), int(win), ctypes.c_void_p(d_cc)))

    def aggregate_dev(self, d_cc, d_w, weights_mode, nx, ny, dmin, dmax, P1, P2, NDIR, MGM, felz, fix, refinement,
                      d_out, d_outcost, d_S=0):
        ri = self.lib.mgmb200_refinement_index(refinement.encode())
        self._check(self.lib.mgmb200_aggregate_dev(self._ctx, ctypes.c_void_p(d_cc), ctypes.c_void_p(d_w or 0),
                                                   int(weights_mode), nx, ny, int(dmin), int(dmax), ctypes.c_float(P1),
                                                   ctypes.c_float(P2), int(NDIR), int(MGM), int(felz), int(fix), ri,
                                                   ctypes.c_void_p(d_out), ctypes.c_void_p(d_outcost),
                                                   ctypes.c_void_p(d_S or 0)))

    def aggregate_sweeps_dev(self, d_cc, d_w, weights_mode, nx, ny, dmin, dmax, P1, P2, NDIR, MGM, felz, sweep_mask):
        self._check(self.lib.mgmb200_aggregate_sweeps_dev(self._ctx, ctypes.c_void_p(d_cc), ctypes.c_void_p(d_w or 0),
                                                          int(weights_mode), nx, ny, int(dmin), int(dmax),
                                                          ctypes.c_float(P1), ctypes.c_float(P2), int(NDIR), int(MGM),
                                                          int(felz), ctypes.c_uint(sweep_mask)))

    def sweep_volume(self, sweep):
        p = ctypes.c_void_p()
        n = ctypes.c_size_t()
        self._check(self.lib.mgmb200_sweep_volume(self._ctx, int(sweep), ctypes.byref(p), ctypes.byref(n)))
        return (p.value or 0), n.value

    def ipc_export(self, d_ptr):
        h = (ctypes.c_ubyte * 64)()
        self._check(self.lib.mgmb200_ipc_export(self._ctx, ctypes.c_void_p(d_ptr), h))
        return bytes(h)

    def ipc_open(self, handle):
        h = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
        p = ctypes.c_void_p()
        self._check(self.lib.mgmb200_ipc_open(self._ctx, h, ctypes.byref(p)))
        return p.value

    def ipc_close(self, d_ptr):
        self._check(self.lib.mgmb200_ipc_close(self._ctx, ctypes.c_void_p(d_ptr)))

    def finish_rows_dev(self, sweep_ptrs, d_cc, nx, ny, dmin, dmax, NDIR, fix, refinement, row_begin, row_end, d_out,
                        d_outcost):
        arr = (ctypes.c_void_p * 16)(*[ctypes.c_void_p(p) for p in list(sweep_ptrs) + [0] * (16 - len(sweep_ptrs))])
        ri = self.lib.mgmb200_refinement_index(refinement.encode())
        self._check(self.lib.mgmb200_finish_rows_dev(self._ctx, arr, ctypes.c_void_p(d_cc), nx, ny, int(dmin), int(dmax),
                                                     int(NDIR), int(fix), ri, int(row_begin), int(row_end),
                                                     ctypes.c_void_p(d_out), ctypes.c_void_p(d_outcost)))

    def aggregate_batch_dev(self, d_ccs, nx, ny, dmin, dmax, P1, P2, NDIR, MGM, felz, fix, refinement, d_outs, d_outcosts,
                            d_ws=None):
        """npairs stereo pairs of one shape in shared launches (mgmb200_aggregate_batch_dev)."""
        n = len(d_ccs)
        mk = lambda ps: (ctypes.c_void_p * n)(*[ctypes.c_void_p(p) for p in ps])
        ri = self.lib.mgmb200_refinement_index(refinement.encode())
        self._check(self.lib.mgmb200_aggregate_batch_dev(self._ctx, n, mk(d_ccs), mk(d_ws) if d_ws else None, nx, ny,
                                                         int(dmin), int(dmax), ctypes.c_float(P1), ctypes.c_float(P2),
                                                         int(NDIR), int(MGM), int(felz), int(fix), ri, mk(d_outs),
                                                         mk(d_outcosts)))

    def sweeps_alloc(self, nx, ny, dmin, dmax, NDIR):
        self._check(self.lib.mgmb200_sweeps_alloc(self._ctx, nx, ny, int(dmin), int(dmax), int(NDIR)))

    def sweeps_release(self):
        self._check(self.lib.mgmb200_sweeps_release(self._ctx))

    def aggregate_sweeps_slabs_dev(self, d_cc, d_w, weights_mode, nx, ny, dmin, dmax, P1, P2, NDIR, MGM, felz, sweep_mask,
                                   nslabs, slab_rows, slab_volumes):
        """slab_volumes[p][r]: device pointer of the volume that receives rows [r*slab_rows,(r+1)*slab_rows) of sweep p."""
        flat = [slab_volumes[p][r] for p in range(NDIR) for r in range(nslabs)]
        arr = (ctypes.c_void_p * len(flat))(*[ctypes.c_void_p(p) for p in flat])
        self._check(self.lib.mgmb200_aggregate_sweeps_slabs_dev(self._ctx, ctypes.c_void_p(d_cc), ctypes.c_void_p(d_w or 0),
                                                                int(weights_mode), nx, ny, int(dmin), int(dmax),
                                                                ctypes.c_float(P1), ctypes.c_float(P2), int(NDIR), int(MGM),
                                                                int(felz), ctypes.c_uint(sweep_mask), int(nslabs),
                                                                int(slab_rows), arr))

    def sum_sweeps_dev(self, nx, ny, dmin, dmax, sweep_mask, d_sum):
        self._check(self.lib.mgmb200_sum_sweeps_dev(self._ctx, nx, ny, int(dmin), int(dmax), ctypes.c_uint(sweep_mask),
                                                    ctypes.c_void_p(d_sum)))

    def finish_sum_dev(self, d_sum, d_cc, nx, ny, dmin, dmax, NDIR, fix, refinement, row_begin, row_end, d_out, d_outcost):
        ri = self.lib.mgmb200_refinement_index(refinement.encode())
        self._check(self.lib.mgmb200_finish_sum_dev(self._ctx, ctypes.c_void_p(d_sum), ctypes.c_void_p(d_cc), nx, ny,
                                                    int(dmin), int(dmax), int(NDIR), int(fix), ri, int(row_begin),
                                                    int(row_end), ctypes.c_void_p(d_out), ctypes.c_void_p(d_outcost)))

    def pad_volume_dev(self, d_dense, d_padded, nx, ny, L, label_major=0):
        self._check(self.lib.mgmb200_pad_volume_dev(self._ctx, ctypes.c_void_p(d_dense), ctypes.c_void_p(d_padded), nx,
                                                    ny, int(L), int(label_major)))

    def unpad_volume_dev(self, d_padded, d_dense, nx, ny, L):
        self._check(self.lib.mgmb200_unpad_volume_dev(self._ctx, ctypes.c_void_p(d_padded), ctypes.c_void_p(d_dense),
                                                      nx, ny, int(L)))
