"""Not a test: round-2 experiment sweep B (one GPU): whole fused step of the headline workload vs options."""
import sys
sys.path.insert(0, ".")
import torch
import mgm_b200
from bench import WORKLOADS, synth_pair

ctx = mgm_b200.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)

def setup(name):
    wl = WORKLOADS[name]
    W, H, L = wl["W"], wl["H"], wl["L"]
    VS = ctx.padded_labels(L)
    u, v = synth_pair(W, H, L, 0)
    du, dv = torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda()
    dcc = torch.empty((H, W, VS), device="cuda")
    ctx.costvolume_dev(du.data_ptr(), dv.data_ptr(), W, H, 1, -(L - 1), 0, "census" if wl["dist"] == "census" else "none",
                       wl["dist"], float("inf"), wl["win"], dcc.data_ptr())
    out = torch.empty((H, W), device="cuda"); cost = torch.empty((H, W), device="cuda")
    ctx.synchronize()
    return wl, dcc, out, cost

def fused(name, opts, state, reps=3):
    wl, dcc, out, cost = state
    W, H, L = wl["W"], wl["H"], wl["L"]
    ctx.set_option("reset")
    for k, v in opts.items():
        ctx.set_option(k, v)
    ts = []
    for i in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            ctx.aggregate_dev(dcc.data_ptr(), 0, 0, W, H, -(L - 1), 0, wl["P1"], wl["P2"], wl["NDIR"], wl["K"], wl["felz"], 1,
                              wl["refine"], out.data_ptr(), cost.data_ptr())
            e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    info = ctx.last_launch_info()
    print("%-44s %-44s %8.3f ms rows=%d/%d launches=%d" % (name[:44], str(opts), min(ts[1:]), info["rows_axis"], info["rows_diag"],
                                                            info["kernel_launches"]), flush=True)

for name in sys.argv[1:] or ["cfg3_2048x1536x256_census3_O8_TSGM3_trunclinear", "cfg2_1920x1080x128_census5_O8_TSGM2"]:
    st = setup(name)
    fused(name, {}, st)
    for o in ({"cc_pf": 0}, {"cc_pf": 2}, {"cc_pf": 3}, {"cc_pf": 5}, {"rows_axis": 48}, {"rows_axis": 40}, {"rows_axis": 40, "rows_diag": 40},
              {"rows_axis": 32}, {"fin_tile": "64x16"}, {"fin_tile": "256x8"}, {"no_fused_finish": 1}):
        fused(name, o, st)
    del st
    torch.cuda.empty_cache()
