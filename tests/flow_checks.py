"""Checks of whole command-line flows against the golden fixtures written by tests/golden/make_golden_flows.py (outputs
of the unmodified reference binary).  The same two functions check the CUDA path (`mgm_b200.Context`, GPU suite) and the
oracle's composition of the stages (`OracleImpl`, CPU suite): both expose stereo_lr / stereo_ranges with the keyword
names of mgm_b200.api.Context."""
import numpy as np

import oracle as O


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


def _kw(p):
    return dict(P1=p["P1"], P2=p["P2"], NDIR=p["NDIR"], MGM=p["K"], use_felzenszwalb_potentials=p["felz"], aP=p["aP"],
                aThresh=p["aThresh"], distance=p["distance"], census_ncc_win=p["win"], refinement=p["refinement"])


def check_flow_lr(g, impl):
    p = g["params"]
    r = impl.stereo_lr(g["u"], g["v"], dmin=p["dmin"], dmax=p["dmax"], testlrrl=1, testlrrl_tau=p["tau"], median=p["median"],
                       want_backproj=True, **_kw(p))
    assert same(r["out_nolr"], g["nolr"]), g["name"]
    assert same(r["out"], g["disp"]) and same(r["outcost"], g["cost"]), g["name"]
    assert same(np.asarray(r["backproj"]).reshape(g["back"].shape), g["back"]), g["name"]
    assert np.isnan(g["disp"]).any() and np.isfinite(g["disp"]).any()   # the left-right test rejected some pixels


def check_flow_ranges(g, impl):
    p = g["params"]
    out, cost, lo, hi = impl.stereo_ranges(g["u"], g["v"], g["lo"], g["hi"], tsgm_iter=p["iters"], truncDist=p["truncDist"],
                                           **_kw(p))
    ok = np.isfinite(g["cost"])   # where no label is finite the reference's label is uninitialised (mgm_core.cc:594)
    assert same(cost, g["cost"]) and same(out[ok], g["disp"][ok]), g["name"]
    assert lo.shape == g["lo"].shape and (hi >= lo).all()


class OracleImpl:
    """The oracle's composition of the flows, behind the keyword interface of mgm_b200.api.Context."""

    @staticmethod
    def _okw(P1, P2, NDIR, MGM, use_felzenszwalb_potentials, aP, aThresh, distance, census_ncc_win, refinement, truncDist=np.inf):
        return dict(P1=P1, P2=P2, NDIR=NDIR, K=MGM, felz=use_felzenszwalb_potentials, aP=aP, aThresh=aThresh,
                    distance=distance, win=census_ncc_win, refinement=refinement, truncDist=truncDist)

    def stereo_lr(self, u, v, dmin=-30, dmax=30, P1=8.0, P2=32.0, NDIR=4, MGM=4, use_felzenszwalb_potentials=0,
                  sgm_fix_overcount=1, aP=1.0, aThresh=5.0, prefilter="none", distance="ad", truncDist=np.inf,
                  census_ncc_win=3, refinement="none", testlrrl=1, testlrrl_tau=1.0, median=0, want_backproj=False,
                  buffers=None):
        kw = self._okw(P1, P2, NDIR, MGM, use_felzenszwalb_potentials, aP, aThresh, distance, census_ncc_win, refinement, truncDist)
        L = O.orc_pipeline(u, v, dmin, dmax, **kw)
        offL = O.orc_median(L["out"], median) if median else L["out"]
        r = dict(out_nolr=offL, out=offL, outcost=L["outcost"])
        if testlrrl:
            R = O.orc_pipeline(v, u, -dmax, -dmin, **kw)
            offR = O.orc_median(R["out"], median) if median else R["out"]
            r.update(out=O.orc_leftright(offL, offR, testlrrl_tau), outR=O.orc_leftright(offR, offL, testlrrl_tau),
                     outcostR=R["outcost"])
        if want_backproj:
            r["backproj"] = O.orc_backproject(r["out"], u, v)
        return r

    def stereo_ranges(self, u, v, dminI, dmaxI, tsgm_iter=1, P1=8.0, P2=32.0, NDIR=4, MGM=4, use_felzenszwalb_potentials=0,
                      sgm_fix_overcount=1, aP=1.0, aThresh=5.0, prefilter="none", distance="ad", truncDist=np.inf,
                      census_ncc_win=3, refinement="none"):
        u = np.asarray(u, np.float32)
        nch = 1 if u.ndim == 2 else u.shape[0]
        lo, hi = np.array(dminI, np.float32), np.array(dmaxI, np.float32)
        emin = int(lo.min()) - 3 * (tsgm_iter - 1)
        emax = int(hi.max()) + 3 * (tsgm_iter - 1)
        w = O.orc_weights(u, aP, aThresh)
        cc = O.orc_costvolume_ranges(u, v, lo, hi, emin, emax, prefilter, distance, truncDist, census_ncc_win)
        slo, shi = lo.copy(), hi.copy()
        o = oc = None
        for _ in range(tsgm_iter):
            rr = O.orc_mgm_ranges(cc, lo, hi, w, emin, slo, shi, np.float32(P1) * nch, np.float32(P2) * nch, NDIR, MGM,
                                  use_felzenszwalb_potentials, sgm_fix_overcount)
            o, oc = O.orc_refine_ranges(rr["S"], slo, shi, emin, rr["out"], rr["outcost"], refinement)
            slo, shi, (gmin, gmax) = O.orc_update_range(o, slo, shi)
            slo[~np.isfinite(slo)] = gmin
            shi[~np.isfinite(shi)] = gmax
        return o, oc, slo, shi
