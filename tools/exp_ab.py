"""Not a test: A/B of library builds (MGMB200_LIBRARY) on the sweeps-only and fused headline step."""
import os, sys
sys.path.insert(0, ".")
import torch
import mgm_b200, mgm_b200.api
mgm_b200.api.exported_symbols = lambda: []      # older builds lack the newest entry points
ctx = mgm_b200.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
tag = os.path.basename(os.environ.get("MGMB200_LIBRARY", "default"))
cases = [(2048, 1536, 256, 3, 1), (1920, 1080, 128, 2, 0)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for (W, H, L, K, felz) in cases:
    VS = ctx.padded_labels(L)
    cc = torch.rand((H, W, VS), device="cuda") * 60
    cc[..., L:] = float("inf")
    out = torch.empty((H, W), device="cuda"); cost = torch.empty((H, W), device="cuda")
    P1, P2 = (2.0, 20000.0) if felz else (8.0, 32.0)
    def t(fn, reps=3):
        ts = []
        for i in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream); fn(); e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return min(ts[1:])
    r = []
    for mask in (0xFF, 0x0F, 0xF0):
        r.append(t(lambda: ctx.aggregate_sweeps_dev(cc.data_ptr(), 0, 0, W, H, -(L - 1), 0, P1, P2, 8, K, felz, mask)))
    f = t(lambda: ctx.aggregate_dev(cc.data_ptr(), 0, 0, W, H, -(L - 1), 0, P1, P2, 8, K, felz, 1, "vfit", out.data_ptr(), cost.data_ptr()))
    print("%-22s %dx%dx%d K%d %s: sweeps all %.3f axis %.3f diag %.3f | fused step %.3f ms" % (tag, W, H, L, K, "trunc" if felz else "sgm", r[0], r[1], r[2], f), flush=True)
    del cc
    torch.cuda.empty_cache()
