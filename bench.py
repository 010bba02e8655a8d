#!/usr/bin/env python
"""bench.py -- MGM hot-path benchmark (contract in the task statement).

One "step" = one pass of the hot path over one synthetic stereo pair of the headline
configuration (BASELINE.json configs[2], the one the metric is quoted on):
2048x1536, 256 disparities, census 3x3, -O 8, TSGM=3, truncated-linear potentials.

  value : Gdisp-updates/s = W*H*L*NDIR / t, t = device time from "cost volume resident in
          HBM" to "disparity + cost maps resident in HBM" (aggregation sweeps + ordered sum +
          over-count fix + WTA + sub-pixel), CUDA events on the launching stream.  That region is ONE
          kernel launch (the aggregation kernel with the finish stage fused in as tile work): `roofline`
          describes it, with the two-launch split (sweeps only / finish only) measured beside it.
  e2e   : same metric through the reference-facing C-ABI call mgmb200_stereo() with HOST
          buffers: H2D of the two images, weights, cost volume, aggregation, refinement, D2H of
          the two maps, all inside the timed region.
  --impl reference : the reference's own CPU implementation (oracle/_ref, compiled from the
          unmodified sources) timed on the host cores on a bounded crop of the same workload.

  e2e_cli_flow : the default command-line flow (both directions, median, left-right tests)
          through mgmb200_stereo_lr with host buffers.

Multi-GPU (torchrun, one rank per GPU): `value` is the batch layout, one independent stereo pair
per GPU and no data-path collective (weak scaling).  `sweep_sharded` reports the north_star
layout beside it: the 8 sweeps of ONE pair sharded over the ranks (sweep p on rank p mod N),
every rank finishes a slab of rows reading the other ranks' sweep volumes over NVLink in sweep
order (bit-identical to 1 GPU), the two maps are all-gathered with NCCL (`--shard sweeps` makes
it the timed step).
"""
import argparse
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
    # name: W, H, L, census win, NDIR, TSGM, trunc-linear, P1, P2, refinement
    "cfg3_2048x1536x256_census3_O8_TSGM3_trunclinear": dict(W=2048, H=1536, L=256, win=3, NDIR=8, K=3, felz=1,
                                                            P1=2.0, P2=20000.0, refine="vfit"),
    "cfg2_1920x1080x128_census5_O8_TSGM2": dict(W=1920, H=1080, L=128, win=5, NDIR=8, K=2, felz=0, P1=8.0, P2=32.0,
                                                refine="vfit"),
    "small_640x480x64_census3_O8_TSGM3_trunclinear": dict(W=640, H=480, L=64, win=3, NDIR=8, K=3, felz=1, P1=2.0,
                                                          P2=20000.0, refine="vfit"),
}
DEFAULT_WORKLOAD = "cfg3_2048x1536x256_census3_O8_TSGM3_trunclinear"


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


def cpu_reference_rate(wl, target_seconds=12.0, seed=0):
    """Times the reference's own mgm() (oracle/_ref, OpenMP, all host threads) on a crop of the workload.
    Returns (Gupd/s, dict)."""
    import oracle as O
    flavour = "_flat" if O.ref_lib("_flat") is not None else ""
    kind = "reference"
    if O.ref_lib(flavour) is None:
        kind = "port"   # oracle/_ref not built on this box: fall back to the C restatement
    L, NDIR, K = wl["L"], wl["NDIR"], wl["K"]
    cores = os.cpu_count() or 1

    def run(cw, ch):
        u, v = synth_pair(cw, ch, L, seed)
        if kind == "reference":
            cc = O.ref_costvolume(u, v, -(L - 1), 0, "census", "census", np.inf, wl["win"], flavour=flavour)
            r = O.ref_mgm(cc, None, -(L - 1), wl["P1"], wl["P2"], NDIR, K, wl["felz"], 1, want_S=False, flavour=flavour)
            t = r["seconds"]
        else:
            cc = O.orc_costvolume(u, v, -(L - 1), 0, "census", "census", np.inf, wl["win"])
            t0 = time.perf_counter()
            O.orc_mgm(cc, None, -(L - 1), wl["P1"], wl["P2"], NDIR, K, wl["felz"], 1, want_S=False)
            t = time.perf_counter() - t0
        return cw * ch * L * NDIR / t / 1e9, t

    # calibrate on a small crop, then size the sample for ~target_seconds
    cw, ch = 256, 96
    rate, t = run(cw, ch)
    scale = max(1.0, min(target_seconds / max(t, 1e-3), wl["W"] * wl["H"] / (cw * ch)))
    ch2 = int(min(wl["H"], max(ch, ch * np.sqrt(scale))))
    cw2 = int(min(wl["W"], max(cw, cw * scale * ch / ch2)))
    rate, t = run(cw2, ch2)
    info = {"value": round(rate, 4), "unit": "Gdisp-updates/s", "cores": cores if kind == "reference" else 1,
            "kind": kind,
            "sample": "%s mgm() on a %dx%dx%d crop of the workload (%.1f s, OMP threads=%s, build flavour '%s')" %
                      ("reference" if kind == "reference" else "oracle C port", cw2, ch2, L, t,
                       os.environ.get("OMP_NUM_THREADS", "all"), flavour or "default")}
    return rate, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rows-per-band", type=int, default=0)
    ap.add_argument("--shard", default="batch", choices=["batch", "sweeps"],
                    help="N>1: 'batch' = one stereo pair per GPU, no data-path collective (weak scaling); "
                         "'sweeps' = one pair, sweeps sharded over the GPUs + ordered peer-memory finish (strong)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    W, H, L, NDIR, K = wl["W"], wl["H"], wl["L"], wl["NDIR"], wl["K"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    updates = W * H * L * NDIR
    config = {"workload": args.workload, "W": W, "H": H, "L": L, "NDIR": NDIR, "TSGM": K,
              "potentials": "truncated-linear" if wl["felz"] else "sgm", "P1": wl["P1"], "P2": wl["P2"],
              "distance": "census %dx%d" % (wl["win"], wl["win"]), "refinement": wl["refine"],
              "l2": "inputs (%.2f GB cost volume) larger than L2" % (W * H * L * 4 / 1e9)}

    if args.impl == "reference":
        if rank != 0:
            return
        if world > 1 or os.environ.get("OMP_NUM_THREADS") == "1":
            os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)   # torchrun pins it to 1; use all host threads
        t0 = time.perf_counter()
        rates = []
        info = None
        for i in range(args.warmup + args.steps):
            r, info = cpu_reference_rate(wl, target_seconds=8.0, seed=i)
            if i >= args.warmup:
                rates.append(r)
            if time.perf_counter() - t0 > 240 and len(rates) >= 1:
                break
        val = float(np.mean(rates))
        info["value"] = round(val, 4)
        line = {"impl": "reference", "metric": "Gdisp-updates/s (W*H*L*Ndirs)", "value": round(val, 4),
                "unit": "Gdisp-updates/s", "n_gpus": args.gpus, "steps": len(rates), "warmup": args.warmup,
                "ms_per_step": None, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config, "cpu_baseline": info,
                "e2e": {"value": round(val, 4), "unit": "Gdisp-updates/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import mgm_b200
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
    batch_mode = world > 1 and args.shard == "batch"
    u, v = synth_pair(W, H, L, seed=rank if batch_mode else 0)
    hu = torch.from_numpy(u).pin_memory()
    hv = torch.from_numpy(v).pin_memory()
    hout = torch.empty((H, W), dtype=torch.float32).pin_memory()
    hcost = torch.empty((H, W), dtype=torch.float32).pin_memory()

    with torch.cuda.stream(stream):
        du = hu.cuda(non_blocking=True)
        dv = hv.cuda(non_blocking=True)
        dcc = torch.empty((H, W, VS), dtype=torch.float32, device="cuda")
        dout = torch.empty((H, W), dtype=torch.float32, device="cuda")
        dcost = torch.empty((H, W), dtype=torch.float32, device="cuda")
        ctx.costvolume_dev(du.data_ptr(), dv.data_ptr(), W, H, 1, dmin, dmax, "census", "census", float("inf"),
                           wl["win"], dcc.data_ptr())
    stream.synchronize()

    from mgm_b200 import sharding
    my_mask = sharding.sweep_mask(NDIR, world, rank)
    peer_ptrs = None
    slabs = sharding.row_slabs(H, world)
    rows = [a for a, _ in slabs] + [H]
    if world > 1:
        # (sweeps mode) one un-timed run so that the sweep volumes exist, then exchange their IPC handles once
        ctx.aggregate_sweeps_dev(dcc.data_ptr(), 0, 0, W, H, dmin, dmax, wl["P1"], wl["P2"], NDIR, K, wl["felz"], my_mask)
        ctx.synchronize()
        peer_ptrs = sharding.exchange_sweep_handles(ctx, dist, NDIR, world, rank)
        gout = [torch.empty((rows[r + 1] - rows[r], W), dtype=torch.float32, device="cuda") for r in range(world)]
        gcost = [torch.empty((rows[r + 1] - rows[r], W), dtype=torch.float32, device="cuda") for r in range(world)]

    def step_device():
        """cost volume resident -> maps resident (the metric's timed region)"""
        if world == 1 or batch_mode:
            ctx.aggregate_dev(dcc.data_ptr(), 0, 0, W, H, dmin, dmax, wl["P1"], wl["P2"], NDIR, K, wl["felz"], 1,
                              wl["refine"], dout.data_ptr(), dcost.data_ptr())
        else:
            ctx.aggregate_sweeps_dev(dcc.data_ptr(), 0, 0, W, H, dmin, dmax, wl["P1"], wl["P2"], NDIR, K, wl["felz"],
                                     my_mask)
            with torch.cuda.stream(stream):
                dist.barrier()   # all sweeps of all ranks are complete before anyone reads peer memory
            ctx.finish_rows_dev(peer_ptrs, dcc.data_ptr(), W, H, dmin, dmax, NDIR, 1, wl["refine"], rows[rank],
                                rows[rank + 1], dout.data_ptr(), dcost.data_ptr())
            with torch.cuda.stream(stream):
                dist.all_gather(gout, dout[rows[rank]:rows[rank + 1]].contiguous())
                dist.all_gather(gcost, dcost[rows[rank]:rows[rank + 1]].contiguous())

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
    ms_dev = timed(step_device, args.steps, args.warmup)
    info = ctx.last_launch_info()
    ms_sweeps = None
    if batch_mode:   # also measure the north_star layout (sweeps sharded, ordered exchange) in the same run
        batch_mode = False
        ms_sweeps = timed(step_device, max(2, args.steps // 2), 2)
        batch_mode = True

    # per-kernel split on rank 0 / single GPU: aggregation kernel vs finish kernel (roofline of the dominant one)
    ms_agg = ms_fin = None
    if world == 1:
        full_mask = (1 << NDIR) - 1
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        sweeps = None
        tot_a = tot_f = 0.0
        for i in range(args.steps):
            with torch.cuda.stream(stream):
                ev[0].record(stream)
                ctx.aggregate_sweeps_dev(dcc.data_ptr(), 0, 0, W, H, dmin, dmax, wl["P1"], wl["P2"], NDIR, K,
                                         wl["felz"], full_mask)
                ev[1].record(stream)
                if sweeps is None:
                    sweeps = [ctx.sweep_volume(p)[0] for p in range(NDIR)]
                ctx.finish_rows_dev(sweeps, dcc.data_ptr(), W, H, dmin, dmax, NDIR, 1, wl["refine"], 0, H,
                                    dout.data_ptr(), dcost.data_ptr())
                ev[2].record(stream)
            torch.cuda.synchronize()
            tot_a += ev[0].elapsed_time(ev[1]); tot_f += ev[1].elapsed_time(ev[2])
        ms_agg, ms_fin = tot_a / args.steps, tot_f / args.steps

    # end-to-end through the reference-facing C-ABI call with host buffers
    ms_e2e = None
    if world == 1 or batch_mode:
        # host buffers are pinned (cudaHostAlloc through torch): the two images in, the two maps out
        pu, pv, pout, pcost = hu.numpy(), hv.numpy(), hout.numpy(), hcost.numpy()

        def step_e2e():
            o, c = ctx.stereo(pu, pv, dmin=dmin, dmax=dmax, P1=wl["P1"], P2=wl["P2"], NDIR=NDIR, MGM=K,
                              use_felzenszwalb_potentials=wl["felz"], distance="census", census_ncc_win=wl["win"],
                              refinement=wl["refine"], out=pout, outcost=pcost)
            return o
        for _ in range(max(1, args.warmup - 1)):
            step_e2e()
        barrier_sync()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        torch.cuda.synchronize()
        ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
        if dist is not None:
            t = torch.tensor([ms_e2e], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t.item())
    # secondary: the default command-line flow (SURVEY 8(d) "CLI-equivalent"): both directions, median of radius 1,
    # left-right tests, all inside one mgmb200_stereo_lr call with host images in and host maps out
    ms_lr = None
    if world == 1:
        lr_bufs = {k: torch.empty((H, W), dtype=torch.float32).pin_memory().numpy()
                   for k in ("out", "outcost", "out_nolr", "outR", "outcostR")}

        def step_lr():
            return ctx.stereo_lr(pu, pv, dmin=dmin, dmax=dmax, P1=wl["P1"], P2=wl["P2"], NDIR=NDIR, MGM=K,
                                 use_felzenszwalb_potentials=wl["felz"], distance="census", census_ncc_win=wl["win"],
                                 refinement=wl["refine"], testlrrl=1, median=1, buffers=lr_bufs)
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
    nunits = world if batch_mode else 1          # stereo pairs processed per step by the whole job
    value = nunits * updates / (ms_dev * 1e-3) / 1e9
    line = {"metric": "Gdisp-updates/s (W*H*L*Ndirs)", "value": round(value, 3), "unit": "Gdisp-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_dev, 4),
            "higher_is_better": True, "scaling": "weak" if (world == 1 or batch_mode) else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "gpu_launches": info["kernel_launches"] * args.steps * world,
            "launch_info": info, "clocks": clocks,
            "mpix_per_s_device": round(nunits * W * H / (ms_dev * 1e-3) / 1e6, 2)}
    sweeps_desc = "sweeps sharded %d per GPU, ordered peer-memory finish over NVLink + NCCL all_gather of the maps" % (
        (NDIR + world - 1) // world)
    if world > 1 and batch_mode:
        config["parallelism"] = "batch: one stereo pair per GPU (%d pairs per step), no data-path collective" % world
        line["sweep_sharded"] = {"value": round(updates / (ms_sweeps * 1e-3) / 1e9, 3), "unit": "Gdisp-updates/s",
                                 "ms_per_step": round(ms_sweeps, 4), "scaling": "strong", "parallelism": sweeps_desc,
                                 "note": "one pair, north_star layout; bit-identical to 1 GPU"}
    elif world > 1:
        config["parallelism"] = sweeps_desc
    if ms_agg is not None:
        agg_bytes = 8.0 * updates   # read C + write the sweep's message, fp32, per label update
        fin_bytes = 4.0 * W * H * L * (NDIR + 1)
        ach = agg_bytes / (ms_agg * 1e-3) / 1e9
        traffic = None
        try:   # DRAM bytes of the same kernel/workload from the committed ncu --set full capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[args.workload]["mgm_aggregate_kernel"]["traffic"]
        except Exception:
            pass
        split = {"aggregation_only": {"kernel": "mgm_aggregate_kernel (sweeps only)", "ms_per_launch": round(ms_agg, 4),
                                      "achieved": round(ach, 1), "frac": round(ach / peak, 4),
                                      "algorithmic_bytes_per_launch": agg_bytes},
                 "finish_only": {"kernel": "mgm_wta_kernel", "ms_per_launch": round(ms_fin, 4),
                                 "achieved": round(fin_bytes / (ms_fin * 1e-3) / 1e9, 1),
                                 "frac": round(fin_bytes / (ms_fin * 1e-3) / 1e9 / peak, 4)}}
        if info["kernel_launches"] == 1:
            # the timed step is ONE launch: the aggregation kernel with the finish stage fused in as tile work
            # (ordered sum + fix + WTA + sub-pixel of tiles whose bands are complete); its algorithmic bytes are the
            # whole step's: 8 B per label update + 4 B * (NDIR + 1) per cell re-read by the finish
            step_bytes = agg_bytes + fin_bytes
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[args.workload]["mgm_aggregate_kernel_fused"]["traffic"]
            except Exception:
                traffic = None
            line["roofline"] = {"bound": "hbm", "kernel": "mgm_aggregate_kernel (sweeps + fused finish tiles)",
                                "achieved": round(step_bytes / (ms_dev * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                                "frac": round(step_bytes / (ms_dev * 1e-3) / 1e9 / peak, 4), "traffic": traffic,
                                "peak_source": peak_src, "ms_per_launch": round(ms_dev, 4),
                                "algorithmic_bytes_per_launch": step_bytes, "unfused_split": split}
        else:
            line["roofline"] = {"bound": "hbm", "kernel": "mgm_aggregate_kernel", "achieved": round(ach, 1),
                                "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": traffic,
                                "peak_source": peak_src, "ms_per_launch": round(ms_agg, 4),
                                "algorithmic_bytes_per_launch": agg_bytes, "finish_kernel": split["finish_only"],
                                "whole_step": {"bytes": 12.0 * updates + 4.0 * W * H * L,
                                               "frac": round((12.0 * updates + 4.0 * W * H * L) / (ms_dev * 1e-3) / 1e9 / peak, 4)}}
    if ms_e2e is not None:
        line["e2e"] = {"value": round(nunits * updates / (ms_e2e * 1e-3) / 1e9, 3), "unit": "Gdisp-updates/s",
                       "h2d_bytes_per_step": nunits * 2 * W * H * 4, "d2h_bytes_per_step": nunits * 2 * W * H * 4,
                       "ms_per_step": round(ms_e2e, 3), "mpix_per_s": round(nunits * W * H / (ms_e2e * 1e-3) / 1e6, 2),
                       "call": "mgmb200_stereo (pinned host images in, pinned host maps out)"}
    if ms_lr is not None:
        line["e2e_cli_flow"] = {"ms_per_pair": round(ms_lr, 3), "mpix_per_s": round(W * H / (ms_lr * 1e-3) / 1e6, 2),
                                "value": round(2 * updates / (ms_lr * 1e-3) / 1e9, 3), "unit": "Gdisp-updates/s",
                                "call": "mgmb200_stereo_lr: L->R + R->L runs, median radius 1, left-right tests "
                                        "(TESTLRRL=1 MEDIAN=1, mgm.cc:372-424); host images in, host maps out"}
    if not args.no_cpu_baseline and world == 1:
        try:
            _, cb = cpu_reference_rate(wl)
            line["cpu_baseline"] = cb
        except Exception as e:   # the checker is optional for the measurement itself
            line["cpu_baseline"] = {"value": None, "error": str(e)}
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
