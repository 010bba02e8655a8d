"""Not a test: lean SGM kernels (aggregate_sgm.cu) vs the generic kernel on the SGM workloads, whole fused step and sweeps only."""
import sys
sys.path.insert(0, ".")
import torch
import mgm_b200

ctx = mgm_b200.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)

def t(fn, reps=3):
    ts = []
    for i in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream); fn(); e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[1:])

import os
FELZ = int(os.environ.get("FELZ", "0"))
cases = [(1920, 1080, 128, 2), (1242, 375, 192, 4), (2048, 1536, 256, 3), (4096, 4096, 64, 2), (640, 480, 64, 2)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
optsets = [{}, {"no_lean_sgm": 1}, {"cc_pf": 0}, {"cc_pf": 6}] if not FELZ else [{}, {"no_lean_trunc": 1}]
P1, P2 = (2.0, 20000.0) if FELZ else (8.0, 32.0)
if os.environ.get("OPTS"):
    import json
    optsets = json.loads(os.environ["OPTS"])
for (W, H, L, K) in cases:
    VS = ctx.padded_labels(L)
    cc = torch.rand((H, W, VS), device="cuda") * 60
    cc[..., L:] = float("inf")
    out = torch.empty((H, W), device="cuda"); cost = torch.empty((H, W), device="cuda")
    for opts in optsets:
        ctx.set_option("reset")
        for k, v in opts.items():
            ctx.set_option(k, v)
        r = [t(lambda: ctx.aggregate_sweeps_dev(cc.data_ptr(), 0, 0, W, H, -(L - 1), 0, P1, P2, 8, K, FELZ, mask)) for mask in (0xFF, 0x0F, 0xF0)]
        f = t(lambda: ctx.aggregate_dev(cc.data_ptr(), 0, 0, W, H, -(L - 1), 0, P1, P2, 8, K, FELZ, 1, "vfit", out.data_ptr(), cost.data_ptr()))
        info = ctx.last_launch_info()
        print("%dx%dx%d K%d felz=%d %-22s sweeps all %.3f axis %.3f diag %.3f | step %.3f ms rows=%d/%d" % (
            W, H, L, K, FELZ, str(opts), r[0], r[1], r[2], f, info["rows_axis"], info["rows_diag"]), flush=True)
    del cc
    torch.cuda.empty_cache()
