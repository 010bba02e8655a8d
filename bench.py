#!/usr/bin/env python
"""bench.py -- MGM hot-path benchmark (contract in the task statement).

One "step" = one pass of the hot path over one batch of synthetic input of the named workload.  The default
workload is the configuration the metric is quoted on (BASELINE.json configs[2]): ONE stereo pair 2048x1536,
256 disparities, census 3x3, -O 8, TSGM=3, truncated-linear potentials.

  value : Gdisp-updates/s = pairs*W*H*L*NDIR / t, t = device time from "cost volumes resident in HBM" to
          "disparity + cost maps resident in HBM" (aggregation sweeps + ordered sum + over-count fix + WTA +
          sub-pixel), CUDA events on the launching stream.  For one pair that region is ONE kernel launch (the
          aggregation kernel with the finish stage fused in as tile work): `roofline` describes it.
  e2e   : the same metric through the reference-facing C-ABI call mgmb200_stereo() with HOST buffers: H2D of the
          images, weights, cost volume, aggregation, refinement, D2H of the two maps, all inside the timed region.
  parity: (N=1, on by default) the UNMODIFIED reference (oracle/_ref, OpenMP, all host threads) run on the SAME
          full-size synthetic pair; WTA labels, aggregated volume, costs and sub-pixel disparities of the CUDA path
          are compared with it.  The same run is the `cpu_baseline`.
  --impl reference : the reference's own mgm() on the host cores, full frame of the same workload per step.

Multi-GPU (torchrun, one rank per GPU): `value` is the batch layout (independent stereo pairs sharded over the
GPUs, no data-path collective); `sweep_sharded` reports the north_star layout beside it (the sweeps of ONE pair
sharded over the ranks), `--shard sweeps` makes it the timed step.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2] -- the configuration the metric is quoted on
    "cfg3_2048x1536x256_census3_O8_TSGM3_trunclinear": dict(W=2048, H=1536, L=256, dist="census", win=3, NDIR=8, K=3,
                                                            felz=1, P1=2.0, P2=20000.0, refine="vfit", pairs=1),
    # configs[1]
    "cfg2_1920x1080x128_census5_O8_TSGM2": dict(W=1920, H=1080, L=128, dist="census", win=5, NDIR=8, K=2, felz=0,
                                                P1=8.0, P2=32.0, refine="vfit", pairs=1),
    # configs[3]: 32 KITTI-shape pairs, command-line defaults otherwise (-t ad, TSGM=4, P1=8, P2=32; mgm.cc:186,303-318)
    "cfg4_32x1242x375x192_ad_O8_TSGM4": dict(W=1242, H=375, L=192, dist="ad", win=3, NDIR=8, K=4, felz=0, P1=8.0,
                                             P2=32.0, refine="none", pairs=32),
    # configs[4]: satellite tile, NCC 5x5, 16 sweeps (sweeps 8-15 are DEFINED by this build)
    "cfg5_4096x4096x64_ncc5_O16_TSGM4": dict(W=4096, H=4096, L=64, dist="ncc", win=5, NDIR=16, K=4, felz=0, P1=8.0,
                                             P2=32.0, refine="none", pairs=1),
    "small_640x480x64_census3_O8_TSGM3_trunclinear": dict(W=640, H=480, L=64, dist="census", win=3, NDIR=8, K=3, felz=1,
                                                          P1=2.0, P2=20000.0, refine="vfit", pairs=1),
}
DEFAULT_WORKLOAD = "cfg3_2048x1536x256_census3_O8_TSGM3_trunclinear"
METRIC = "Gdisp-updates/s (W*H*L*Ndirs)"
UNIT = "Gdisp-updates/s"


def synth_pair(W, H, L, seed=0):
    """right image = 3-octave box-blurred noise quantised to 0..255, left image = right
    image warped by d(x,y) = -round(L/4 + L/8 sin(0.01x) cos(0.013y)) plus integer noise in [-2,2]."""
    rng = np.random.default_rng(seed)
    base = rng.random((H, W + L)).astype(np.float64)
    acc = base.copy()
    for k in (2, 4, 8):
        cs = np.cumsum(np.pad(base, ((0, 0), (k, 0))), axis=1)
        acc += (cs[:, k:] - cs[:, :-k]) / k
    acc = np.floor(255 * (acc - acc.min()) / (acc.max() - acc.min() + 1e-9))
    v = acc[:, :W].astype(np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    d = -np.round(L / 4 + (L / 8) * np.sin(0.01 * xx) * np.cos(0.013 * yy)).astype(np.int64)
    xs = np.clip(xx + d, 0, W - 1)
    u = (v[yy, xs] + rng.integers(-2, 3, (H, W))).astype(np.float32)
    return np.ascontiguousarray(u), np.ascontiguousarray(v)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([s.strip() for s in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, reasons, smmax = [], set(), 0
        for s in self.samples:
            try:
                sm.append(float(s[0])); smmax = max(smmax, float(s[1]))
                for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smmax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_source_sha():
    """Identifies the kernel sources an ncu traffic capture belongs to (profiles/traffic.json is refused when stale)."""
    h = hashlib.sha256()
    for f in ("aggregate.cu", "aggregate.cuh", "aggregate_dev.cuh", "aggregate_sgm.cu", "aggregate_trunc.cu", "aggregate_sgmw.cu", "aggregate_plan.cu",
              "wta_device.cuh", "wta.cuh", "common.cuh"):
        h.update(open(os.path.join(ROOT, "mgm_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def recorded_traffic(workload, key):
    """DRAM bytes per launch of the named kernel from the committed `ncu --set full` capture, or (None, why)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = t[workload][key]
        if e.get("kernel_source_sha") != kernel_source_sha():
            return None, "profiles/traffic.json was captured on other kernel sources (%s): refused" % e.get("kernel_source_sha")
        return e["traffic"], "ncu dram__bytes_read.sum + dram__bytes_write.sum, %s" % e.get("capture", "profiles/")
    except Exception as ex:
        return None, "no capture recorded (%s)" % type(ex).__name__


def config_of(name, wl):
    W, H, L = wl["W"], wl["H"], wl["L"]
    dist = wl["dist"] + (" %dx%d" % (wl["win"], wl["win"]) if wl["dist"] in ("census", "ncc") else "")
    return {"workload": name, "pairs": wl["pairs"], "W": W, "H": H, "L": L, "NDIR": wl["NDIR"], "TSGM": wl["K"],
            "potentials": "truncated-linear" if wl["felz"] else "sgm", "P1": wl["P1"], "P2": wl["P2"],
            "distance": dist, "refinement": wl["refine"],
            "l2": "inputs (%.2f GB of cost volumes per step) larger than L2" % (wl["pairs"] * W * H * L * 4 / 1e9)}


# ------------------------------------------------------------------------------------------ the reference on the CPU
class ReferenceRun:
    """The unmodified reference (oracle/_ref) on one synthetic pair of the workload; falls back to the C port
    (oracle/mgm_oracle.c, one thread) when oracle/_ref is absent or the workload is outside the reference's
    defined range (NDIR > 8)."""

    def __init__(self, wl, seed=0, crop=None):
        import oracle as O
        self.O, self.wl = O, wl
        self.flavour = "_flat" if O.ref_lib("_flat") is not None else ""
        self.kind = "reference" if (O.ref_lib(self.flavour) is not None and wl["NDIR"] <= 8) else "port"
        self.W, self.H = crop if crop else (wl["W"], wl["H"])
        self.L = wl["L"]
        self.dmin, self.dmax = -(self.L - 1), 0
        self.u, self.v = synth_pair(self.W, self.H, self.L, seed)
        self.cores = (os.cpu_count() or 1) if self.kind == "reference" else 1
        self.cc = None
        self.t_cc = None

    def costvolume(self):
        if self.cc is None:
            O, wl = self.O, self.wl
            pf = "census" if wl["dist"] == "census" else "none"
            t0 = time.perf_counter()
            if self.kind == "reference":
                self.cc = O.ref_costvolume(self.u, self.v, self.dmin, self.dmax, pf, wl["dist"], np.inf, wl["win"],
                                           flavour=self.flavour)
            else:
                self.cc = O.orc_costvolume(self.u, self.v, self.dmin, self.dmax, pf, wl["dist"], np.inf, wl["win"])
            self.t_cc = time.perf_counter() - t0
        return self.cc

    def mgm(self, want_S):
        """-> dict(out, outcost, S, seconds): mgm() of the reference (mgm_core.cc:408-613)"""
        O, wl = self.O, self.wl
        cc = self.costvolume()
        if self.kind == "reference":
            return O.ref_mgm(cc, None, self.dmin, wl["P1"], wl["P2"], wl["NDIR"], wl["K"], wl["felz"], 1, want_S=want_S,
                             flavour=self.flavour)
        t0 = time.perf_counter()
        r = O.orc_mgm(cc, None, self.dmin, wl["P1"], wl["P2"], wl["NDIR"], wl["K"], wl["felz"], 1, want_S=want_S)
        r["seconds"] = time.perf_counter() - t0
        return r

    def refine(self, r):
        O, wl = self.O, self.wl
        if wl["refine"] == "none":
            return r["out"], r["outcost"]
        if self.kind == "reference":
            return O.ref_refine(r["S"], self.dmin, r["out"], r["outcost"], wl["refine"], flavour=self.flavour)
        return O.orc_refine(r["S"], self.dmin, r["out"], r["outcost"], wl["refine"])

    def rate(self, seconds):
        return self.W * self.H * self.L * self.wl["NDIR"] / seconds / 1e9

    def describe(self, seconds):
        what = ("unmodified reference mgm() (oracle/_ref, build flavour '%s', OpenMP)" % (self.flavour or "default")
                if self.kind == "reference" else "oracle C port orc_mgm() (one thread)")
        full = (self.W, self.H) == (self.wl["W"], self.wl["H"])
        return "%s on %s %dx%dx%d pair of the workload, %d sweeps: %.2f s (cost volume %.2f s not counted), OMP threads=%s" % (
            what, "the FULL" if full else "a cropped", self.W, self.H, self.L, self.wl["NDIR"], seconds, self.t_cc or 0.0,
            os.environ.get("OMP_NUM_THREADS", "all"))


def reference_crop(wl):
    """The reference runs the full frame when it is defined for the workload and the frame costs about a minute or
    less on 16 cores; otherwise (NDIR=16: C port, one thread) a bounded crop."""
    if wl["NDIR"] <= 8:
        return None
    return (512, 384)


def rel_err(a, b):
    """max |a-b| / |b| over finite b != 0 (a, b float32 arrays), ignoring positions where both are non-finite"""
    a = a.astype(np.float64).ravel(); b = b.astype(np.float64).ravel()
    ok = np.isfinite(b) & (b != 0)
    bad_nonfinite = int(np.count_nonzero(np.isfinite(a) != np.isfinite(b)))
    if not ok.any():
        return 0.0, bad_nonfinite
    return float(np.max(np.abs(a[ok] - b[ok]) / np.abs(b[ok]))), bad_nonfinite


def parity_check(ctx, wl, ref, torch):
    """The CUDA path against the reference on the same pair.  Returns the `parity` object and the reference's
    mgm() seconds."""
    W, H, L, dmin, dmax = ref.W, ref.H, ref.L, ref.dmin, ref.dmax
    r = ref.mgm(want_S=True)
    ro, rc = ref.refine(r)
    kw = dict(dmin=dmin, dmax=dmax, P1=wl["P1"], P2=wl["P2"], NDIR=wl["NDIR"], MGM=wl["K"],
              use_felzenszwalb_potentials=wl["felz"], distance=wl["dist"], census_ncc_win=wl["win"])
    go, gc = ctx.stereo(ref.u, ref.v, refinement="none", **kw)          # WTA labels + their costs
    gro, grc = ctx.stereo(ref.u, ref.v, refinement=wl["refine"], **kw)  # refined
    wta_mismatch = int(np.count_nonzero(~((go == r["out"]) | (np.isnan(go) & np.isnan(r["out"])))))
    cost_rel, cost_nf = rel_err(gc, r["outcost"])
    sub_rel, sub_nf = rel_err(gro, ro)
    subc_rel, _ = rel_err(grc, rc)
    par = {"against": ref.kind, "shape": "%dx%dx%d" % (W, H, L), "sweeps": wl["NDIR"], "wta_mismatch": wta_mismatch,
           "cost_max_rel": cost_rel, "subpixel_max_rel": sub_rel, "subpixel_cost_max_rel": subc_rel,
           "nonfinite_mismatch": cost_nf + sub_nf,
           # bit comparison of the refined maps; NaNs (vfit on a flat minimum divides 0/0, refine.h:70-92) count as equal
           # whatever their payload (x86 produces the negative quiet NaN, the GPU the positive one)
           "disparity_bit_mismatch": int(np.count_nonzero((gro.view(np.int32) != ro.view(np.int32)) & ~(np.isnan(gro) & np.isnan(ro)))),
           "cost_bit_mismatch": int(np.count_nonzero((grc.view(np.int32) != rc.view(np.int32)) & ~(np.isnan(grc) & np.isnan(rc)))),
           "ref_seconds": round(r["seconds"], 3), "ref_costvolume_seconds": round(ref.t_cc or 0.0, 3)}
    # aggregated volume S (after the over-count fix), compared on the device chunk by chunk
    try:
        VS = ctx.padded_labels(L)
        du = torch.from_numpy(ref.u).cuda(); dv = torch.from_numpy(ref.v).cuda()
        dcc = torch.empty((H, W, VS), dtype=torch.float32, device="cuda")
        dS = torch.empty((H, W, L), dtype=torch.float32, device="cuda")   # dense [pix][L], like the reference's return value
        do = torch.empty((H, W), dtype=torch.float32, device="cuda"); dc = torch.empty_like(do)
        ctx.set_stream(0)
        ctx.costvolume_dev(du.data_ptr(), dv.data_ptr(), W, H, 1, dmin, dmax, "census" if wl["dist"] == "census" else "none",
                           wl["dist"], float("inf"), wl["win"], dcc.data_ptr())
        ctx.aggregate_dev(dcc.data_ptr(), 0, 0, W, H, dmin, dmax, wl["P1"], wl["P2"], wl["NDIR"], wl["K"], wl["felz"], 1,
                          "none", do.data_ptr(), dc.data_ptr(), dS.data_ptr())
        ctx.synchronize()
        bit_mis, max_rel, cc_mis = 0, 0.0, 0
        rows = max(1, (256 << 20) // (W * L * 4))
        for y0 in range(0, H, rows):
            y1 = min(H, y0 + rows)
            rs = torch.from_numpy(r["S"][y0:y1]).cuda()
            gs = dS[y0:y1]
            same = (gs.view(torch.int32) == rs.view(torch.int32)) | (torch.isnan(gs) & torch.isnan(rs))
            bit_mis += int((~same).sum().item())
            fin = torch.isfinite(rs) & (rs != 0)
            if bool(fin.any()):
                max_rel = max(max_rel, float(((gs[fin].double() - rs[fin].double()).abs() / rs[fin].double().abs()).max().item()))
            rcc = torch.from_numpy(ref.cc[y0:y1]).cuda()
            cc_mis += int((dcc[y0:y1, :, :L].view(torch.int32) != rcc.view(torch.int32)).sum().item())
        par["S_bit_mismatch"] = bit_mis
        par["S_max_rel"] = max_rel
        par["costvolume_bit_mismatch"] = cc_mis
        del du, dv, dcc, dS, do, dc
        torch.cuda.empty_cache()
    except Exception as e:   # pragma: no cover - reported, never hidden
        par["S_error"] = str(e)[:200]
    par["ok"] = bool(wta_mismatch == 0 and cost_rel <= 1e-4 and sub_rel <= 1e-4 and par["nonfinite_mismatch"] == 0 and
                     par.get("S_max_rel", 0.0) <= 1e-4)
    return par, r["seconds"]


def run_reference_arm(args, name, wl):
    """--impl reference: the reference's mgm() on the host cores, the full frame of the workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    if world > 1 or os.environ.get("OMP_NUM_THREADS") == "1":
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)   # torchrun pins it to 1; use all host threads
    ref = ReferenceRun(wl, seed=0, crop=reference_crop(wl))
    ref.costvolume()
    t0 = time.perf_counter()
    secs = []
    budget = 300.0
    for i in range(args.warmup + args.steps):
        r = ref.mgm(want_S=False)
        if i >= args.warmup:
            secs.append(r["seconds"])
        el = time.perf_counter() - t0
        per = el / (i + 1)
        # bounded run: skip the remaining warm-up steps / stop early when the budget would be exceeded
        if i < args.warmup and el + per * (args.steps + args.warmup - i - 1) > budget:
            args.warmup = i + 1
        if secs and el + per > budget:
            break
    if not secs:
        secs = [r["seconds"]]
    sec = float(np.mean(secs))
    val = ref.rate(sec)
    info = {"value": round(val, 4), "unit": UNIT, "cores": ref.cores, "kind": ref.kind, "sample": ref.describe(sec)}
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(secs), "warmup": args.warmup, "ms_per_step": round(sec * 1e3 * wl["pairs"], 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(name, wl), "cpu_baseline": info,
            "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ the CUDA arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the reference run (parity + cpu_baseline)")
    ap.add_argument("--no-parity", action="store_true", help="alias of --no-cpu-baseline")
    ap.add_argument("--rows-per-band", type=int, default=0)
    ap.add_argument("--shard", default="batch", choices=["batch", "sweeps"],
                    help="N>1: 'batch' = independent pairs sharded over the GPUs, no data-path collective; "
                         "'sweeps' = one pair, its sweeps sharded over the GPUs (north_star layout)")
    ap.add_argument("--exchange", default="ordered", choices=["ordered", "allreduce"],
                    help="sweeps layout: 'ordered' = sweep-ordered sum from slab-distributed volumes (bit-exact); "
                         "'allreduce' = NCCL all-reduce of the per-GPU partial sums")
    args = ap.parse_args()
    name = args.workload
    wl = WORKLOADS[name]
    if args.impl == "reference":
        return run_reference_arm(args, name, wl)

    W, H, L, NDIR, K, pairs = wl["W"], wl["H"], wl["L"], wl["NDIR"], wl["K"], wl["pairs"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    updates = W * H * L * NDIR          # per pair
    config = config_of(name, wl)

    import torch
    import mgm_b200
    from mgm_b200 import sharding
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = mgm_b200.Context(local_rank)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    if args.rows_per_band:
        ctx.set_rows_per_band(args.rows_per_band)

    dmin, dmax = -(L - 1), 0
    VS = ctx.padded_labels(L)
    pf = "census" if wl["dist"] == "census" else "none"
    batch_mode = world == 1 or args.shard == "batch"
    # batch layout: a single-pair workload runs one pair per GPU (weak scaling); a multi-pair workload splits its
    # pairs over the ranks (the whole job's work is fixed: strong scaling)
    if pairs == 1:
        my_pairs = [rank] if (world > 1 and batch_mode) else [0]
        job_pairs = world if (world > 1 and batch_mode) else 1
        scaling = "weak" if batch_mode else "strong"
    else:
        my_pairs = list(range(rank, pairs, world)) if batch_mode else list(range(pairs))
        job_pairs = pairs
        scaling = "strong" if world > 1 else "weak"
    npair = len(my_pairs)

    hu, hv, dcc = [], [], []
    for s in my_pairs:
        u, v = synth_pair(W, H, L, seed=s)
        hu.append(torch.from_numpy(u).pin_memory()); hv.append(torch.from_numpy(v).pin_memory())
    hout = torch.empty((H, W), dtype=torch.float32).pin_memory()
    hcost = torch.empty((H, W), dtype=torch.float32).pin_memory()
    with torch.cuda.stream(stream):
        for i in range(npair):
            du = hu[i].cuda(non_blocking=True); dv = hv[i].cuda(non_blocking=True)
            c = torch.empty((H, W, VS), dtype=torch.float32, device="cuda")
            ctx.costvolume_dev(du.data_ptr(), dv.data_ptr(), W, H, 1, dmin, dmax, pf, wl["dist"], float("inf"),
                               wl["win"], c.data_ptr())
            dcc.append(c)
        dout = [torch.empty((H, W), dtype=torch.float32, device="cuda") for _ in range(npair)]
        dcost = [torch.empty((H, W), dtype=torch.float32, device="cuda") for _ in range(npair)]
    stream.synchronize()

    agg_args = (W, H, dmin, dmax, wl["P1"], wl["P2"], NDIR, K, wl["felz"])

    def step_batch():
        """cost volumes resident -> maps resident (the metric's timed region), this rank's pairs"""
        if npair > 1 and hasattr(ctx, "aggregate_batch_dev"):
            ctx.aggregate_batch_dev([c.data_ptr() for c in dcc], *agg_args, 1, wl["refine"],
                                    [o.data_ptr() for o in dout], [o.data_ptr() for o in dcost])
        else:
            for i in range(npair):
                ctx.aggregate_dev(dcc[i].data_ptr(), 0, 0, *agg_args, 1, wl["refine"], dout[i].data_ptr(), dcost[i].data_ptr())

    # north_star layout (ONE pair -- seed 0 on every rank --, its sweeps sharded over the ranks)
    sw = None
    if world > 1:
        if my_pairs[0] == 0:
            dcc_sw = dcc[0]
        else:
            u0, v0 = synth_pair(W, H, L, seed=0)
            with torch.cuda.stream(stream):
                du = torch.from_numpy(u0).cuda(); dv = torch.from_numpy(v0).cuda()
                dcc_sw = torch.empty((H, W, VS), dtype=torch.float32, device="cuda")
                ctx.costvolume_dev(du.data_ptr(), dv.data_ptr(), W, H, 1, dmin, dmax, pf, wl["dist"], float("inf"),
                                   wl["win"], dcc_sw.data_ptr())
            stream.synchronize()
        sw_out = torch.empty((H, W), dtype=torch.float32, device="cuda")
        sw_cost = torch.empty((H, W), dtype=torch.float32, device="cuda")
        sw = sharding.SweepSharded(ctx, dist, torch, stream, W, H, dmin, dmax, wl["P1"], wl["P2"], NDIR, K, wl["felz"],
                                   wl["refine"], world, rank)
        sw.setup(dcc_sw)

    def step_sweeps():
        sw.step(dcc_sw, sw_out, sw_cost, exchange=args.exchange)

    def barrier_sync():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier_sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        barrier_sync()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev = timed(step_batch if batch_mode else step_sweeps, args.steps, args.warmup)
    info = ctx.last_launch_info()
    launches_per_step = info["kernel_launches"] * (npair if not hasattr(ctx, "aggregate_batch_dev") or npair == 1 else 1)

    sweeps_report = None
    if world > 1:
        # both exchanges of the north_star layout, measured in the same run, with mismatch counts against the
        # single-GPU maps of the same pair (computed on this rank by the batch path)
        one_out = torch.empty((H, W), dtype=torch.float32, device="cuda")
        one_cost = torch.empty((H, W), dtype=torch.float32, device="cuda")
        ctx.aggregate_dev(dcc_sw.data_ptr(), 0, 0, *agg_args, 1, wl["refine"], one_out.data_ptr(), one_cost.data_ptr())
        ctx.synchronize()
        sweeps_report = {}
        for ex in ("ordered", "allreduce"):
            try:
                ms = timed(lambda: sw.step(dcc_sw, sw_out, sw_cost, exchange=ex), max(2, args.steps // 2), 2)
                torch.cuda.synchronize()
                neq = lambda a, b: int((~((a == b) | (torch.isnan(a) & torch.isnan(b)))).sum().item())
                # differences of a whole label or more (sub-pixel refinement aside); NaN (no finite label / flat vfit) = NaN
                wta = lambda a, b: int((((a - b).abs() >= 0.5) | (torch.isnan(a) != torch.isnan(b))).sum().item())
                sweeps_report[ex] = {"value": round(updates / (ms * 1e-3) / 1e9, 3), "unit": UNIT, "ms_per_step": round(ms, 4),
                                     "scaling": "strong", "disparity_mismatch_vs_1gpu": neq(sw_out, one_out),
                                     "label_flips_vs_1gpu": wta(sw_out, one_out),
                                     "cost_mismatch_vs_1gpu": neq(sw_cost, one_cost), "pixels": W * H,
                                     "parallelism": sw.describe(ex)}
            except Exception as e:
                sweeps_report[ex] = {"error": str(e)[:300]}

    # per-kernel split on a single GPU: aggregation kernel vs finish kernel (explains the fused launch)
    ms_agg = ms_fin = None
    if world == 1 and NDIR <= 8:
        full_mask = (1 << NDIR) - 1
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        sweeps = None
        tot_a = tot_f = 0.0
        nrep = min(args.steps, 3)
        for i in range(nrep):
            with torch.cuda.stream(stream):
                ev[0].record(stream)
                ctx.aggregate_sweeps_dev(dcc[0].data_ptr(), 0, 0, *agg_args, full_mask)
                ev[1].record(stream)
                if sweeps is None:
                    sweeps = [ctx.sweep_volume(p)[0] for p in range(NDIR)]
                ctx.finish_rows_dev(sweeps, dcc[0].data_ptr(), W, H, dmin, dmax, NDIR, 1, wl["refine"], 0, H,
                                    dout[0].data_ptr(), dcost[0].data_ptr())
                ev[2].record(stream)
            torch.cuda.synchronize()
            tot_a += ev[0].elapsed_time(ev[1]); tot_f += ev[1].elapsed_time(ev[2])
        ms_agg, ms_fin = tot_a / nrep, tot_f / nrep

    # end-to-end through the reference-facing C-ABI call with host buffers (pinned): images in, maps out
    ms_e2e = None
    if batch_mode:
        pout, pcost = hout.numpy(), hcost.numpy()

        kw_e2e = dict(dmin=dmin, dmax=dmax, P1=wl["P1"], P2=wl["P2"], NDIR=NDIR, MGM=K, use_felzenszwalb_potentials=wl["felz"],
                      distance=wl["dist"], census_ncc_win=wl["win"], refinement=wl["refine"])
        if npair > 1:
            pouts = [torch.empty((H, W), dtype=torch.float32).pin_memory().numpy() for _ in range(npair)]
            pcosts = [torch.empty((H, W), dtype=torch.float32).pin_memory().numpy() for _ in range(npair)]

        def step_e2e():
            if npair > 1:   # one call for the rank's pairs: mgmb200_stereo_batch
                ctx.stereo_batch([h.numpy() for h in hu], [h.numpy() for h in hv], outs=pouts, outcosts=pcosts, **kw_e2e)
            else:
                ctx.stereo(hu[0].numpy(), hv[0].numpy(), out=pout, outcost=pcost, **kw_e2e)
        for _ in range(max(1, args.warmup - 1) if npair == 1 else 1):
            step_e2e()
        barrier_sync()
        nst = args.steps if npair == 1 else max(1, min(args.steps, 2))
        t0 = time.perf_counter()
        for _ in range(nst):
            step_e2e()
        torch.cuda.synchronize()
        ms_e2e = (time.perf_counter() - t0) * 1e3 / nst
        if dist is not None:
            t = torch.tensor([ms_e2e], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t.item())
    # secondary: the default command-line flow (SURVEY 8(d) "CLI-equivalent"): both directions, median of radius 1,
    # left-right tests, all inside one mgmb200_stereo_lr call with host images in and host maps out
    ms_lr = None
    if world == 1 and pairs == 1 and NDIR <= 8:
        lr_bufs = {k: torch.empty((H, W), dtype=torch.float32).pin_memory().numpy()
                   for k in ("out", "outcost", "out_nolr", "outR", "outcostR")}

        def step_lr():
            return ctx.stereo_lr(hu[0].numpy(), hv[0].numpy(), dmin=dmin, dmax=dmax, P1=wl["P1"], P2=wl["P2"], NDIR=NDIR,
                                 MGM=K, use_felzenszwalb_potentials=wl["felz"], distance=wl["dist"],
                                 census_ncc_win=wl["win"], refinement=wl["refine"], testlrrl=1, median=1, buffers=lr_bufs)
        step_lr()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(min(args.steps, 3)):
            step_lr()
        torch.cuda.synchronize()
        ms_lr = (time.perf_counter() - t0) * 1e3 / min(args.steps, 3)
    clocks = sampler.finish() if rank == 0 else None

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    units = job_pairs if batch_mode else 1        # stereo pairs processed per step by the whole job
    value = units * updates / (ms_dev * 1e-3) / 1e9
    line = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_dev, 4), "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "gpu_launches": launches_per_step * args.steps * world, "launch_info": info, "clocks": clocks,
            "mpix_per_s_device": round(units * W * H / (ms_dev * 1e-3) / 1e6, 2)}
    if world > 1 and batch_mode:
        config["parallelism"] = ("batch: %d independent stereo pairs per step over %d GPUs (%d per GPU), no data-path "
                                 "collective" % (job_pairs, world, npair))
        line["sweep_sharded"] = sweeps_report
    elif world > 1:
        config["parallelism"] = sw.describe(args.exchange)
        line["sweep_sharded"] = sweeps_report

    # roofline of the dominant kernel: the aggregation launch (sweeps + finish tiles fused in).  Algorithmic bytes per
    # launch: 8 B per label update (read C, write the sweep's message) + 4 B * (NDIR + 1) per cell for the ordered
    # re-read by the finish; measured peak from MEASURED_PEAKS.json.  With N GPUs in the batch layout every
    # GPU runs the same launches on its own pairs: bytes and time are per GPU.
    agg_bytes = 8.0 * updates
    fin_bytes = 4.0 * W * H * L * (NDIR + 1)
    if batch_mode:
        step_bytes = (agg_bytes + fin_bytes) * npair
        fused = info["kernel_launches"] == 1
        traffic, traffic_src = recorded_traffic(name, "mgm_aggregate_kernel_fused" if fused else "mgm_aggregate_kernel")
        # which aggregation kernel the library launches for this workload (csrc/aggregate.cu agg_launch)
        kname = "mgm_aggregate_kernel" if NDIR > 8 else ("mgm_aggregate_trunc_kernel" if wl["felz"] else "mgm_aggregate_sgm_kernel")
        roof = {"bound": "hbm", "kernel": kname + " (sweeps + fused finish tiles)" if fused else
                kname + " + mgm_wta_kernel",
                "achieved": round(step_bytes / (ms_dev * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                "frac": round(step_bytes / (ms_dev * 1e-3) / 1e9 / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "ms_per_launch": round(ms_dev / max(1, launches_per_step), 4),
                "launches_per_step": launches_per_step,
                "algorithmic_bytes_per_launch": step_bytes / max(1, launches_per_step), "per": "GPU"}
        if ms_agg is not None:
            ach = agg_bytes / (ms_agg * 1e-3) / 1e9
            roof["unfused_split"] = {
                "aggregation_only": {"kernel": kname + " (sweeps only)", "ms_per_launch": round(ms_agg, 4),
                                     "achieved": round(ach, 1), "frac": round(ach / peak, 4),
                                     "algorithmic_bytes_per_launch": agg_bytes},
                "finish_only": {"kernel": "mgm_wta_kernel", "ms_per_launch": round(ms_fin, 4),
                                "achieved": round(fin_bytes / (ms_fin * 1e-3) / 1e9, 1),
                                "frac": round(fin_bytes / (ms_fin * 1e-3) / 1e9 / peak, 4)}}
        line["roofline"] = roof
    else:
        # sweeps layout: per GPU NDIR/N sweeps (8 B per update) + the finish of H/N rows
        per_gpu = agg_bytes / world + fin_bytes / world
        line["roofline"] = {"bound": "hbm", "kernel": "mgm_aggregate_kernel (this GPU's sweeps) + mgm_wta_kernel (its slab)",
                            "achieved": round(per_gpu / (ms_dev * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                            "frac": round(per_gpu / (ms_dev * 1e-3) / 1e9 / peak, 4), "traffic": None,
                            "peak_source": peak_src, "per": "GPU",
                            "floor": "one axis sweep is a dependency chain of maxii + maxjj pixel steps (DESIGN.md section 5)"}
    if ms_e2e is not None:
        line["e2e"] = {"value": round(units * updates / (ms_e2e * 1e-3) / 1e9, 3), "unit": UNIT,
                       "h2d_bytes_per_step": units * 2 * W * H * 4, "d2h_bytes_per_step": units * 2 * W * H * 4,
                       "ms_per_step": round(ms_e2e, 3), "mpix_per_s": round(units * W * H / (ms_e2e * 1e-3) / 1e6, 2),
                       "call": ("mgmb200_stereo_batch" if npair > 1 else "mgmb200_stereo") + " (pinned host images in, pinned host maps out)"}
    if ms_lr is not None:
        line["e2e_cli_flow"] = {"ms_per_pair": round(ms_lr, 3), "mpix_per_s": round(W * H / (ms_lr * 1e-3) / 1e6, 2),
                                "value": round(2 * updates / (ms_lr * 1e-3) / 1e9, 3), "unit": UNIT,
                                "call": "mgmb200_stereo_lr: L->R + R->L runs, median radius 1, left-right tests "
                                        "(TESTLRRL=1 MEDIAN=1, mgm.cc:372-424); host images in, host maps out"}
    if not (args.no_cpu_baseline or args.no_parity) and world == 1:
        # parity gate + CPU baseline: ONE run of the reference on the same pair (full frame when defined)
        try:
            ref = ReferenceRun(wl, seed=0, crop=reference_crop(wl))
            par, sec = parity_check(ctx, wl, ref, torch)
            line["parity"] = par
            line["cpu_baseline"] = {"value": round(ref.rate(sec), 4), "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
                                    "sample": ref.describe(sec)}
        except Exception as e:   # reported, never hidden
            line["parity"] = {"ok": False, "error": str(e)[:300]}
            line["cpu_baseline"] = {"value": None, "error": str(e)[:300]}
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
